"""Host-side MJCF flattener (csrc/mjcf_flatten.cpp, replaces xml_parser.h + DynamicModel::LoadModel):
error reporting on files outside the supported model class, tolerance to the reference file's quirks."""
import os
import re

import numpy as np

from conftest import ROOT


def _load(harness, text, tmp_path, name="m.xml"):
    p = tmp_path / name
    p.write_text(text)
    rc = harness.L.hh_load(str(p).encode())
    return rc, harness.L.hh_error().decode()


def test_flatten_errors(harness, oracle, tmp_path):
    good = open(oracle.default_model_path()).read()
    rc, err = _load(harness, "<mujoco><worldbody/></mujoco>", tmp_path)
    assert rc != 0 and "13 joints" in err
    rc, err = _load(harness, "<notmujoco/>", tmp_path)
    assert rc != 0 and "mujoco" in err
    rc, err = _load(harness, good.replace('solver="PGS"', 'solver="Newton"'), tmp_path)
    assert rc != 0 and "PGS" in err
    rc, err = _load(harness, good.replace("<mujoco", "<mujoco><!-- unterminated", 1), tmp_path)
    assert rc != 0 and "comment" in err
    rc = harness.L.hh_load(b"/nonexistent.xml")
    assert rc != 0 and "cannot open" in harness.L.hh_error().decode()
    # and the good file still loads afterwards (the fixture is shared)
    assert harness.L.hh_load(oracle.default_model_path().encode()) == 0


def test_flatten_tolerates_reference_comment_style(harness, oracle, tmp_path):
    """The reference MJCF closes some comments with '--->' (cassie2d_stiff.xml:69,73): not well-formed XML."""
    good = open(oracle.default_model_path()).read()
    txt = good.replace("<worldbody>", "<worldbody>\n  <!-- a comment closed the reference's way --->", 1)
    rc, err = _load(harness, txt, tmp_path)
    assert rc == 0, err
    assert abs(harness.L.hh_total_mass() - 32.822) < 1e-9
    assert harness.L.hh_load(oracle.default_model_path().encode()) == 0
