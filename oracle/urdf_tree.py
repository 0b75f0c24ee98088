"""ORACLE (test infrastructure, NOT product code): URDF + STL -> Bullet-style 33-link tree tables.

Restates, in float64 numpy, what PyBullet's URDF importer does for the one robot the reference
loads (`plen_bullet/src/plen_bullet/plen_env.py:312-315` -> `plen_bullet/src/plen.urdf:504-1489`,
foot meshes `plen_ros/meshes_bin/{r,l}foot.stl` named at `plen.urdf:1097` and `:1263`).

PyBullet/Bullet3 is a third-party, un-vendored, un-pinned dependency of the reference (SURVEY.md
section 8c) so every importer behaviour below is restated from the published algorithm ([RECALL] in
SURVEY.md Appendix A) and exposed as a named constant:

* fixed joints are KEPT as 0-DoF links (no URDF_MERGE_FIXED_LINKS); link index = DFS pre-order with
  children in joint-file order (agrees with the indices printed at `plen_env.py:718-742`);
* URDF <inertia> tensors are ignored; the diagonal inertia is the box inertia of the AABB of the
  link's compound collision shape (child AABB + child margin, + compound margin);
* collision margin 0.001 m for URDF shapes; convex hulls are inflated by it, boxes keep their size.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` leg may
import this module.  The product loader lives in `plen_ml_walk_b200/urdf_loader.py` and folds fixed
links into 19 bodies; this one deliberately keeps all 33 so the two are independent restatements.
"""
from __future__ import annotations

import json
import os
import struct
import xml.etree.ElementTree as ET

import numpy as np

# ---- [RECALL] importer constants -------------------------------
URDF_COLLISION_MARGIN = 0.001      # gUrdfDefaultCollisionMargin, set on every child shape
COMPOUND_MARGIN = 0.001            # margin set on the per-link btCompoundShape (added to its AABB)
CONTACT_BREAKING_FACTOR = 0.02     # gContactBreakingThreshold; threshold = factor * angularMotionDisc

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_JSON = os.path.join(HERE, "data", "plen_tree33.json")

# plen_env.py:318-320 -- Bullet joint indices of the 18 actuated joints, in action/observation order
MOVING_JOINTS = [5, 6, 7, 9, 10, 11, 13, 14, 15, 17, 18, 19, 20, 21, 24, 26, 27, 30]
RIGHT_FOOT_LINK = 11               # plen_env.py:446-456, :784
LEFT_FOOT_LINK = 19                # plen_env.py:457-467, :774


def rpy_to_matrix(rpy):
    """URDF fixed-axis roll/pitch/yaw -> rotation matrix R = Rz(y) Ry(p) Rx(r)."""
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = np.cos(r), np.sin(r), np.cos(p), np.sin(p), np.cos(y), np.sin(y)
    return np.array([
        [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
        [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
        [-sp, cp * sr, cp * cr],
    ])


def _vec(s, n=3):
    v = [float(x) for x in s.split()]
    assert len(v) == n, s
    return np.array(v)


def read_binary_stl_vertices(path):
    """Unique vertices of a binary STL (80-byte header, uint32 count, 50-byte facets)."""
    with open(path, "rb") as f:
        data = f.read()
    (ntri,) = struct.unpack_from("<I", data, 80)
    assert len(data) == 84 + 50 * ntri, "not a binary STL: %s" % path
    tri = np.frombuffer(data, dtype=np.dtype([("n", "<f4", 3), ("v", "<f4", (3, 3)), ("a", "<u2")]),
                        count=ntri, offset=84)
    verts = tri["v"].reshape(-1, 3).astype(np.float64)
    return np.unique(verts, axis=0)


def sole_corner_points(verts_link):
    """Four sole vertices spanning the largest quadrilateral of the flat sole of a foot hull.

    verts_link: [V,3] hull vertices in the LINK frame.  The sole is the set of vertices within
    1e-6 m of the minimum link-frame z (32 coplanar vertices for both PLEN feet);
    its outline is an elongated octagon.  Bullet keeps at most 4 points per persistent manifold and
    its reduction heuristic keeps the subset of largest area ([RECALL], SURVEY.md Appendix A.C1);
    the from-scratch contact generator uses that limiting 4-point set directly: the 4-subset of
    sole vertices with maximum quadrilateral area (ties -> first in angular order), returned
    counter-clockwise seen from above, starting from the (+x,+y) quadrant.
    """
    import itertools
    zmin = verts_link[:, 2].min()
    sole = verts_link[np.abs(verts_link[:, 2] - zmin) < 1e-6]
    c = sole[:, :2].mean(0)
    ang = np.arctan2(sole[:, 1] - c[1], sole[:, 0] - c[0])
    sole = sole[np.argsort(ang)]                       # counter-clockwise polygon order
    x, y = sole[:, 0], sole[:, 1]
    best, best_idx = -1.0, None
    for idx in itertools.combinations(range(len(sole)), 4):
        i = list(idx)
        area = 0.5 * abs(sum(x[i[k]] * y[i[(k + 1) % 4]] - x[i[(k + 1) % 4]] * y[i[k]] for k in range(4)))
        if area > best + 1e-15:
            best, best_idx = area, i
    pts = sole[best_idx]
    a = np.arctan2(pts[:, 1] - c[1], pts[:, 0] - c[0])
    start = int(np.argmin(np.where(a >= 0, a, a + 2 * np.pi)))
    pts = np.roll(pts, -start, axis=0)
    return pts, sole


def hull_vertices(verts_link):
    """Vertices of the convex hull of a foot mesh (link frame), in the order of the STL's unique-vertex list: the support-vertex
    search of the persistent-manifold experiment (plen_oracle_config.manifold_mode) walks them as btConvexHullShape would walk
    its point list after optimizeConvexHull ([RECALL]; only the first-maximum tie break depends on the order)."""
    from scipy.spatial import ConvexHull
    idx = np.sort(ConvexHull(verts_link).vertices)
    return verts_link[idx]


def build_tree(urdf_path, mesh_dir, inertia_from_urdf=False):
    root = ET.parse(urdf_path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    children = {}
    child_names = set()
    for j in joints:
        children.setdefault(j.find("parent").get("link"), []).append(j)
        child_names.add(j.find("child").get("link"))
    base_name = [n for n in links if n not in child_names]
    assert base_name == ["torso"], base_name
    base_name = base_name[0]

    order = []  # (joint element, parent index) in DFS pre-order, children in file order

    def dfs(link_name, parent_idx):
        for j in children.get(link_name, []):
            idx = len(order)
            order.append((j, parent_idx))
            dfs(j.find("child").get("link"), idx)

    dfs(base_name, -1)
    n = len(order)

    def link_props(name):
        l = links[name]
        inertial = l.find("inertial")
        mass = float(inertial.find("mass").get("value"))
        io = inertial.find("origin")
        com = _vec(io.get("xyz"))
        assert np.allclose(_vec(io.get("rpy")), 0.0), "inertial rpy != 0 not needed for PLEN"
        cols = l.findall("collision")
        assert len(cols) == 1
        co = cols[0].find("origin")
        c_xyz, c_R = _vec(co.get("xyz")), rpy_to_matrix(_vec(co.get("rpy")))
        geom = cols[0].find("geometry")[0]
        hull, box = None, None
        if geom.tag == "box":
            half = _vec(geom.get("size")) * 0.5          # box keeps its size (margin is inside)
            box = dict(center=c_xyz.tolist(), rot=c_R.tolist(), half=half.tolist())
            ext = np.abs(c_R) @ half                     # AABB half extents in the link frame
            lo, hi = c_xyz - ext, c_xyz + ext
        elif geom.tag == "mesh":
            fn = geom.get("filename").split("/")[-1]
            v = read_binary_stl_vertices(os.path.join(mesh_dir, fn)) * _vec(geom.get("scale"))
            hull = v @ c_R.T + c_xyz                     # hull vertices, link frame
            lo = hull.min(0) - URDF_COLLISION_MARGIN     # convex hull AABB includes its margin
            hi = hull.max(0) + URDF_COLLISION_MARGIN
        else:
            raise ValueError(geom.tag)
        lo, hi = lo - COMPOUND_MARGIN, hi + COMPOUND_MARGIN
        lx, ly, lz = hi - lo
        inertia = mass / 12.0 * np.array([ly * ly + lz * lz, lx * lx + lz * lz, lx * lx + ly * ly])
        if inertia_from_urdf:      # the URDF_USE_INERTIA_FROM_FILE behaviour (NOT what the reference's loadURDF call asks for):
            it = inertial.find("inertia")      # only for the sensitivity table (scripts/sensitivity_table.py)
            assert all(abs(float(it.get(k))) < 1e-12 for k in ("ixy", "ixz", "iyz")), "off-diagonal URDF inertia"
            inertia = np.array([float(it.get(k)) for k in ("ixx", "iyy", "izz")])
        # compound AABB expressed in the inertial frame (origin = com) for the breaking threshold
        centre = 0.5 * (lo + hi) - com
        radius = 0.5 * np.linalg.norm(hi - lo)
        return dict(name=name, mass=mass, com=com, inertia=inertia, hull=hull, box=box,
                    angular_motion_disc=float(np.linalg.norm(centre) + radius))

    base = link_props(base_name)
    tree = dict(
        n_links=n, base_name=base_name, base_mass=base["mass"], base_com=base["com"].tolist(),
        base_inertia=base["inertia"].tolist(),
        link_names=[], joint_names=[], parent=[], jtype=[], axis=[], R_pj=[], p_pj=[], com=[],
        mass=[], inertia=[], lower=[], upper=[],
    )
    props = []
    for j, parent_idx in order:
        name = j.find("child").get("link")
        pr = link_props(name)
        props.append(pr)
        o = j.find("origin")
        tree["link_names"].append(name)
        tree["joint_names"].append(j.get("name"))
        tree["parent"].append(parent_idx)
        jt = j.get("type")
        assert jt in ("fixed", "revolute"), jt
        tree["jtype"].append(1 if jt == "revolute" else 0)
        ax = _vec(j.find("axis").get("xyz")) if jt == "revolute" else np.zeros(3)
        tree["axis"].append(ax.tolist())
        tree["R_pj"].append(rpy_to_matrix(_vec(o.get("rpy"))).tolist())
        tree["p_pj"].append(_vec(o.get("xyz")).tolist())
        tree["com"].append(pr["com"].tolist())
        tree["mass"].append(pr["mass"])
        tree["inertia"].append(pr["inertia"].tolist())
        lim = j.find("limit")
        tree["lower"].append(float(lim.get("lower")) if lim is not None else 0.0)
        tree["upper"].append(float(lim.get("upper")) if lim is not None else 0.0)

    moving = [i for i in range(n) if tree["jtype"][i] == 1]
    assert moving == sorted(MOVING_JOINTS), moving     # plen_env.py:318-320
    tree["moving_joints"] = MOVING_JOINTS
    feet = []
    for link in (RIGHT_FOOT_LINK, LEFT_FOOT_LINK):
        pr = props[link]
        assert pr["hull"] is not None, pr["name"]
        pts, sole = sole_corner_points(pr["hull"])
        feet.append(dict(link=link, name=pr["name"], points=pts.tolist(), n_sole_vertices=int(len(sole)),
                         breaking_threshold=CONTACT_BREAKING_FACTOR * pr["angular_motion_disc"],
                         hull=hull_vertices(pr["hull"]).tolist()))
    tree["feet"] = feet
    # box colliders of every link but the feet, Bullet link index (-1 = base), for the link-vs-ground contacts
    boxes = [dict(link=-1, name=base_name, **base["box"])]
    for i, pr in enumerate(props):
        if pr["box"] is not None:
            boxes.append(dict(link=i, name=pr["name"], **pr["box"]))
    tree["boxes"] = boxes
    tree["total_mass"] = float(base["mass"] + sum(tree["mass"]))
    return tree


def save_tree(tree, path=DEFAULT_JSON):
    os.makedirs(os.path.dirname(path), exist_ok=True)
    with open(path, "w") as f:
        json.dump(tree, f, indent=1)


def load_tree(path=DEFAULT_JSON):
    with open(path) as f:
        return json.load(f)


if __name__ == "__main__":  # regenerate oracle/data/plen_tree33.json from the read-only reference
    import sys
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    t = build_tree(os.path.join(ref, "plen_bullet/src/plen.urdf"), os.path.join(ref, "plen_ros/meshes_bin"))
    save_tree(t)
    save_tree(build_tree(os.path.join(ref, "plen_bullet/src/plen.urdf"), os.path.join(ref, "plen_ros/meshes_bin"), inertia_from_urdf=True),
              os.path.join(HERE, "data", "plen_tree33_urdf_inertia.json"))
    print("links", t["n_links"], "mass", t["total_mass"], "feet", [(f["name"], f["breaking_threshold"]) for f in t["feet"]])
