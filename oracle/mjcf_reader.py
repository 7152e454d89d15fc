"""TEST INFRASTRUCTURE (oracle side) -- independent MJCF reader for the CPU oracle.

Reads a Cassie MJCF (the reference's `model/cassie2d_stiff.xml`, which is not well-formed
XML: comments end in '--->', cassie2d_stiff.xml:69,73; or the cleaned copy shipped in
`cassierl_b200/model/`) with Python's ElementTree and applies MuJoCo's documented compile
rules [EXT: MuJoCo XML reference] needed by this model class:

  * <compiler angle='degree'>: hinge ref/range in degrees -> radians (cassie2d_stiff.xml:3)
  * <default> joint/geom/motor attributes (cassie2d_stiff.xml:14-18)
  * body frame from pos + xyaxes (x normalised, y Gram-Schmidt, z = x cross y)
  * capsule `fromto` -> (pos = midpoint, z-axis = to-from, half length)
  * hinge default axis (0,0,1), joint pos default (0,0,0)

It is deliberately a different code path from the product's C++ flattener
(`cassierl_b200/csrc/mjcf_flatten.cpp`) so that a parsing mistake in either shows up as a
parity failure.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg
may import this module.
"""
import re
import xml.etree.ElementTree as ET
import numpy as np

GEOM_PLANE, GEOM_SPHERE, GEOM_CAPSULE = 0, 2, 3
JNT_FREE, JNT_SLIDE, JNT_HINGE = 0, 2, 3


def _vec(s, n=None):
    v = np.array([float(x) for x in s.split()], dtype=np.float64)
    if n is not None:
        assert v.size == n, (s, n)
    return v


def _frame_from_xyaxes(xy):
    x = xy[:3] / np.linalg.norm(xy[:3])
    y = xy[3:] - x * np.dot(x, xy[3:])
    y = y / np.linalg.norm(y)
    z = np.cross(x, y)
    return np.stack([x, y, z], axis=1)  # columns = body axes in parent coords


def _frame_from_zaxis(zdir):
    """Rotation whose z column is `zdir` (minimal rotation from (0,0,1)), as MuJoCo's
    fromto handling does with mju_quatZ2Vec [EXT]."""
    z = zdir / np.linalg.norm(zdir)
    k = np.array([0.0, 0.0, 1.0])
    ax = np.cross(k, z)
    s = np.linalg.norm(ax)
    c = float(np.dot(k, z))
    if s < 1e-14:
        return np.eye(3) if c > 0 else np.diag([1.0, -1.0, -1.0])
    ax = ax / s
    K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
    return np.eye(3) + s * K + (1 - c) * (K @ K)


def read_mjcf(path):
    txt = open(path).read()
    txt = re.sub(r"<!--.*?-->", "", txt, flags=re.S)
    root = ET.fromstring(txt)

    comp = root.find("compiler")
    degree = (comp is None) or (comp.get("angle", "degree") == "degree")
    ang = np.pi / 180.0 if degree else 1.0

    opt = root.find("option")
    option = dict(timestep=0.002, iterations=100, tolerance=1e-8, impratio=1.0,
                  gravity=np.array([0, 0, -9.81]), cone="pyramidal", solver="Newton")
    if opt is not None:
        for k in ("timestep", "tolerance", "impratio"):
            if opt.get(k) is not None:
                option[k] = float(opt.get(k))
        if opt.get("iterations") is not None:
            option["iterations"] = int(opt.get("iterations"))
        if opt.get("gravity") is not None:
            option["gravity"] = _vec(opt.get("gravity"), 3)
        for k in ("cone", "solver"):
            if opt.get(k) is not None:
                option[k] = opt.get(k)

    dflt = {"joint": {}, "geom": {}, "motor": {}}
    d = root.find("default")
    if d is not None:
        for k in dflt:
            e = d.find(k)
            if e is not None:
                dflt[k] = dict(e.attrib)

    def attr(el, name, default=None):
        v = el.get(name)
        if v is None:
            v = dflt.get(el.tag, {}).get(name)
        return default if v is None else v

    bodies, joints, geoms, sites = [], [], [], []
    names = {"world": 0}
    bodies.append(dict(name="world", parent=0, pos=np.zeros(3), mat=np.eye(3),
                       ipos=np.zeros(3), mass=0.0, inertia=np.zeros(6)))

    def add_geom(g, bid):
        gtype = attr(g, "type", "sphere")
        if gtype == "mesh":
            return
        size = _vec(attr(g, "size", "0"))
        rec = dict(body=bid, contype=int(attr(g, "contype", "1")),
                   conaffinity=int(attr(g, "conaffinity", "1")),
                   condim=int(attr(g, "condim", "3")),
                   friction=_vec(attr(g, "friction", "1 0.005 0.0001"), 3),
                   solref=_vec(attr(g, "solref", "0.02 1"), 2),
                   solimp=_vec(attr(g, "solimp", "0.9 0.95 0.001")),
                   margin=float(attr(g, "margin", "0")), gap=float(attr(g, "gap", "0")))
        if gtype == "plane":
            rec.update(type=GEOM_PLANE, pos=_vec(attr(g, "pos", "0 0 0"), 3), mat=np.eye(3),
                       size=np.array([0.0, 0.0]))
        elif gtype == "sphere":
            rec.update(type=GEOM_SPHERE, pos=_vec(attr(g, "pos", "0 0 0"), 3), mat=np.eye(3),
                       size=np.array([size[0], 0.0]))
        elif gtype == "capsule":
            ft = _vec(g.get("fromto"), 6)
            a, b = ft[:3], ft[3:]
            rec.update(type=GEOM_CAPSULE, pos=0.5 * (a + b), mat=_frame_from_zaxis(b - a),
                       size=np.array([size[0], 0.5 * np.linalg.norm(b - a)]))
        else:
            raise ValueError("unsupported geom type " + gtype)
        geoms.append(rec)

    def walk(el, parent):
        for ch in el:
            if ch.tag == "geom" and parent == 0 and el.tag == "worldbody":
                add_geom(ch, 0)
        for b in el.findall("body"):
            bid = len(bodies)
            names[b.get("name")] = bid
            mat = np.eye(3)
            if b.get("xyaxes") is not None:
                mat = _frame_from_xyaxes(_vec(b.get("xyaxes"), 6))
            rec = dict(name=b.get("name"), parent=parent, pos=_vec(b.get("pos", "0 0 0"), 3),
                       mat=mat, ipos=np.zeros(3), mass=0.0, inertia=np.zeros(6))
            ine = b.find("inertial")
            if ine is not None:
                rec["ipos"] = _vec(ine.get("pos"), 3)
                rec["mass"] = float(ine.get("mass"))
                rec["inertia"] = _vec(ine.get("fullinertia"), 6)  # Ixx Iyy Izz Ixy Ixz Iyz
            bodies.append(rec)
            for j in b.findall("joint"):
                jt = attr(j, "type", "hinge")
                is_hinge = jt == "hinge"
                limited = attr(j, "limited", "false") == "true"
                rng = _vec(attr(j, "range", "0 0"), 2)
                ref = float(attr(j, "ref", "0"))
                if jt not in ("hinge", "slide", "free"):
                    raise ValueError("unsupported joint type " + jt)
                joints.append(dict(
                    name=j.get("name"), body=bid, type=JNT_FREE if jt == "free" else (JNT_HINGE if is_hinge else JNT_SLIDE),
                    axis=_vec(attr(j, "axis", "0 0 1"), 3), pos=_vec(attr(j, "pos", "0 0 0"), 3),
                    ref=ref * (ang if is_hinge else 1.0), limited=limited,
                    range=rng * (ang if is_hinge else 1.0),
                    damping=float(attr(j, "damping", "0")), armature=float(attr(j, "armature", "0")),
                    stiffness=float(attr(j, "stiffness", "0")),
                    solref=_vec(attr(j, "solreflimit", "0.02 1"), 2),
                    solimp=_vec(attr(j, "solimplimit", "0.9 0.95 0.001"))))
            for g in b.findall("geom"):
                add_geom(g, bid)
            for s in b.findall("site"):
                sites.append(dict(name=s.get("name"), body=bid, pos=_vec(s.get("pos", "0 0 0"), 3)))
            walk(b, bid)

    walk(root.find("worldbody"), 0)

    jnames = {j["name"]: i for i, j in enumerate(joints)}
    eqs = []
    eq = root.find("equality")
    if eq is not None:
        for c in eq.findall("connect"):
            eqs.append(dict(body1=names[c.get("body1")], body2=names[c.get("body2")],
                            anchor=_vec(c.get("anchor"), 3),
                            solref=_vec(c.get("solref", "0.02 1"), 2),
                            solimp=_vec(c.get("solimp", "0.9 0.95 0.001"))))
    acts = []
    ac = root.find("actuator")
    if ac is not None:
        for m in ac.findall("motor"):
            acts.append(dict(joint=jnames[m.get("joint")], gear=float(m.get("gear", "1")),
                             ctrllimited=attr(m, "ctrllimited", "false") == "true",
                             ctrlrange=_vec(m.get("ctrlrange", "0 0"), 2)))
    return dict(option=option, bodies=bodies, joints=joints, geoms=geoms, sites=sites,
                equalities=eqs, actuators=acts, body_names=names, joint_names=jnames)
