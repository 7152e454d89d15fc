// Flattened, planar constants of a Cassie-2D MJCF model: what the device step kernel reads
// from __constant__ memory.  Produced on the host by mjcf_flatten.cpp (replaces the
// reference's XML_Parser::parse_xml_model + DynamicModel::LoadModel,
// CassieRL/cassierl src/xml_parser.h:104-363, src/DynamicModel.cpp:23-235, and MuJoCo's own
// mj_loadXML of the same file, src/Cassie2d/Cassie2d.cpp:48).
//
// Topology is compile-time (it fixes the register layout of the kernel):
//   link 0 = pelvis (dofs 0,1 slide x/z, dof 2 pitch); per leg L (0 left, 1 right) five
//   hinged links: 0 thigh, 1 knee(+shin,+knee_spring), 2 tarsus(+heel_spring), 3 toe, 4 rod.
//   dof(L, j) = 3 + 5 L + j.  Parents: thigh->pelvis, knee->thigh, tarsus->knee, toe->tarsus,
//   rod->thigh.  One loop closure per leg: rod anchor <-> tarsus(heel spring) anchor.
// All vectors are (x, z) pairs in the link's ZERO frame, i.e. the world-aligned frame the
// link has at qpos0; at run time  world = Rot(alpha_link) * v  with alpha about +y.
#pragma once

namespace cassie {

constexpr int kNV = 13;       // RobotInterface.h:53-55 nQ
constexpr int kNU = 6;        // RobotInterface.h:56 nU
constexpr int kLegLinks = 5;
constexpr int kNumCaps = 8;   // per leg: thigh, shin, tarsus, toe capsules (canonical order)
constexpr int kNumSites = 6;  // imu, body_center, L front, L rear, R front, R rear
constexpr int kMaxRows = 46;  // 4 connect + 8 limit + 2*17 contact rows
constexpr int kNumContactSlots = 17;

// link ids inside a leg
enum { kThigh = 0, kKnee = 1, kTarsus = 2, kToe = 3, kRod = 4 };

template <typename T>
struct PlanarModel {
  // pelvis
  T pel_org[2];        // pelvis pivot at qpos0 (world)
  T pel_ref[3];        // ref of the three base joints
  T pel_com[2], pel_mass, pel_inertia;
  // leg links [leg][link]
  T off[2][kLegLinks][2];   // pivot offset in the PARENT's zero frame (from the parent's pivot)
  T sgn[2][kLegLinks];      // +1 / -1: hinge axis = sgn * world +y
  T ang0[2][kLegLinks];     // alpha = alpha_parent + sgn*q + ang0   (ang0 = -sgn*ref)
  T com[2][kLegLinks][2];   // com offset from the pivot, link zero frame
  T mass[2][kLegLinks], inertia[2][kLegLinks];  // composite of welded bodies, about the com
  // per dof
  T damping[kNV], armature[kNV];
  T lim_lo[kNV], lim_hi[kNV];      // joint range (rad); has_limit flag below
  T lim_diag[kNV];                 // dof_invweight0
  int has_limit[kNV];
  T lim_solref[2], lim_solimp[5];  // same for all joints in these models (checked at flatten)
  // actuators
  int act_dof[kNU];
  T act_gear[kNU], act_lo[kNU], act_hi[kNU];
  // loop closures
  T eq_a1[2][2];   // anchor on the rod link, rod zero frame
  T eq_a2[2][2];   // anchor on the tarsus link (heel spring), tarsus zero frame
  T eq_solref[2], eq_solimp[5], eq_diag[2];  // eq_diag = invweight0 tran(rod) + tran(heel_spring)
  // collision geoms vs the floor plane z = 0
  T sph_c[2], sph_r, sph_diag;                     // pelvis sphere
  int cap_link[kNumCaps];                          // leg*8+link ... stored as leg-local link id
  T cap_to[kNumCaps][2], cap_from[kNumCaps][2];    // end points in the link zero frame ('to' first)
  T cap_r[kNumCaps], cap_diag[kNumCaps];           // radius, invweight0 tran of the geom's body
  T con_solref[2], con_solimp[5], con_mu;          // floor-contact parameters (mixed)
  // sites (link -1 = pelvis, else leg*5+link) and offsets
  int site_link[kNumSites];
  T site_off[kNumSites][2];
  // options
  T timestep, gravity_z, tolerance, meaninertia, impratio;
  int iterations;
  T total_mass;
};

// canonical contact slot of capsule c (0..7), end e (0 = 'to', 1 = 'from'): 1 + 2c + e; slot 0
// is the pelvis sphere.  Bit (2*g + e) of the public contact mask uses g = 1 for the pelvis
// sphere and g = 2 + c for capsule c (the geom order of the MJCF file).
constexpr int contact_slot(int cap, int end) { return 1 + 2 * cap + end; }

}  // namespace cassie
