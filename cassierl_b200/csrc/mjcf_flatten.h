// Host-side MJCF reader + planar flattener.
// Replaces the reference's two loaders of the same file: XML_Parser::parse_xml_model +
// DynamicModel::LoadModel (CassieRL/cassierl src/xml_parser.h:104-363,
// src/DynamicModel.cpp:23-235; the controller's RBDL model) and MuJoCo's mj_loadXML
// (src/Cassie2d/Cassie2d.cpp:48; the physics model).  Both are emitted as PlanarModel
// constant blocks: `phys` (MuJoCo compile semantics) and `ctrl` (the RBDL loader's semantics,
// including its xyaxes/ref "hack", DynamicModel.cpp:84-103).
#pragma once
#include <string>
#include "planar_model.h"
#include "tree_model.h"

namespace cassie {

struct FlatModels {
  PlanarModel<double> phys;
  PlanarModel<double> ctrl;
};

// Returns false and fills `err` when the file cannot be read, is not a Cassie-2D-class
// planar model, or violates an assumption of the planar engine.
bool flatten_mjcf_file(const std::string& path, FlatModels* out, std::string* err);
bool flatten_mjcf_text(const std::string& xml, FlatModels* out, std::string* err);

// Free-base 3-D models (cassie3d_stiff.xml) -> TreeModel (tree_model.h).  Same error contract.
bool flatten_tree_file(const std::string& path, tree::TreeModel<double>* out, std::string* err);

template <typename T>
PlanarModel<T> cast_model(const PlanarModel<double>& s);

}  // namespace cassie
