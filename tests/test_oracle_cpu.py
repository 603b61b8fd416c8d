"""CPU tests of the oracle (oracle/splat_oracle.c): pinned against the golden vectors produced by the unmodified
reference on a B200, plus closed-form and structural checks, plus the product's per-surfel math executed on the
host (tests/hostemu) against the oracle."""
import ctypes as C
import os

import numpy as np
import pytest

import util
from util import bits, rel_err
from eggfusion_b200 import synthetic as syn
from oracle import oracle as orc

TOL = 1e-4


# ------------------------------------------------------------------------------------ golden pins (the reference)
@pytest.mark.parametrize("name", util.case_names())
def test_oracle_matches_reference_golden(name):
    path = util.golden_path(name)
    assert os.path.exists(path), "golden vectors missing: run tests/golden/make_golden.py on a GPU box"
    G = np.load(path)
    cam, sc, g, bg, mask, deg, f, b = util.oracle_run(name)
    # index artefacts: bit-exact
    assert f["num_rendered"] == int(G["num_rendered"]) and f["tile_num"] == int(G["tile_num"])
    assert np.array_equal(f["radii"], G["radii"])
    assert np.array_equal(f["active_mask"], G["active_mask"].astype(np.uint8))
    assert np.array_equal(f["tiles_touched"], G["tiles_touched"])
    assert np.array_equal(f["point_list_keys"], G["point_list_keys"])
    assert np.array_equal(f["point_list"], G["point_list"])
    assert np.array_equal(f["ranges"], G["ranges"])
    assert np.array_equal(f["tile_indices"], G["tile_indices"])
    vis = G["vis_index"]
    assert np.array_equal(np.nonzero(f["radii"] > 0)[0], vis)
    # the quantities that decide the index artefacts: bit-exact
    for k in ("means2D", "depths", "cov3D", "normal"):
        assert np.array_equal(bits(f[k][vis]), bits(G[k])), k
    assert np.array_equal(bits(f["conic_opacity"][vis]), bits(G["conic_opacity"]))
    assert np.array_equal(f["clamped"][vis].astype(bool), G["clamped"].astype(bool))
    # the rest: within tolerance (in practice ~1e-7)
    assert rel_err(f["rgb"][vis], G["rgb"]) <= 1e-6
    assert rel_err(f["Jinv"][vis], G["Jinv"]) <= 1e-5
    assert rel_err(f["color"], G["color"]) <= TOL
    assert rel_err(f["out_normal"], G["normal_img"]) <= TOL
    assert rel_err(f["depth"], G["depth"]) <= TOL
    assert rel_err(f["opacity"], G["opacity"]) <= TOL
    gx = (cam.width + 15) // 16
    act = np.zeros((cam.height, cam.width), bool)
    for t in G["tile_indices"][:int(G["tile_num"])]:
        ty, tx = divmod(int(t), gx)
        act[ty * 16:ty * 16 + 16, tx * 16:tx * 16 + 16] = True
    act = act.reshape(-1)
    assert np.mean(f["n_contrib"][act] != G["n_contrib"][act]) <= 2e-4
    assert rel_err(f["final_T"][act], G["final_T"][act]) <= TOL
    assert rel_err(f["final_D"][act], G["final_D"][act]) <= TOL
    for mine, theirs in (("dL_dmean2D", "dL_dmeans2D"), ("dL_dcolors", "dL_dcolors"), ("dL_dopacity", "dL_dopacity"),
                         ("dL_dmeans3D", "dL_dmeans3D"), ("dL_dcov3D", "dL_dcov3D"), ("dL_dsh", "dL_dsh"),
                         ("dL_dscales", "dL_dscales"), ("dL_drotations", "dL_drotations")):
        assert rel_err(b[mine].reshape(-1), G[theirs].reshape(-1)) <= TOL, mine


# ------------------------------------------------------------------------------------ closed form
def test_single_fronto_parallel_surfel_closed_form():
    """One surfel on the optical axis facing the camera: radius, conic, centre alpha, depth and normal by hand."""
    W = H = 64
    cam = syn.Camera(W, H, 50.0, 50.0, 32.0, 32.0)   # integer principal point -> the surfel centre is a pixel centre
    z, s, o = 2.0, 0.12, 0.8
    xyz = np.array([[0, 0, z]], np.float32)
    q = np.array([[0, 1, 0, 0]], np.float32)          # 180 deg about x: normal (0,0,-1) faces the camera
    scales = np.array([[s, s, 0]], np.float32)
    sh = np.zeros((1, 1, 3), np.float32)
    sh[0, 0] = (np.array([0.2, 0.5, 0.9]) - 0.5) / syn.SH_C0
    oc = orc.cam_from_synthetic(cam, 0, 1)
    f = orc.forward(oc, xyz, scales, q, np.array([[o]], np.float32), sh)
    sig2 = (s * cam.fx / z) ** 2 + 0.3
    assert f["radii"][0] == int(np.ceil(3 * np.sqrt(sig2)))
    np.testing.assert_allclose(f["means2D"][0], [32.0, 32.0], atol=1e-4)
    np.testing.assert_allclose(f["conic_opacity"][0], [1 / sig2, 0, 1 / sig2, o], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(f["normal"][0], [0, 0, -1], atol=1e-6)
    np.testing.assert_allclose(f["rgb"][0], [0.2, 0.5, 0.9], atol=1e-6)
    # centre pixel: alpha = o, depth image = z, normal image = alpha * n, colour = alpha * rgb (black background)
    np.testing.assert_allclose(f["color"][:, 32, 32], o * np.array([0.2, 0.5, 0.9]), rtol=1e-5)
    np.testing.assert_allclose(f["depth"][0, 32, 32], z, rtol=1e-5)
    np.testing.assert_allclose(f["opacity"][0, 32, 32], o, rtol=1e-5)
    np.testing.assert_allclose(f["out_normal"][:, 32, 32], [0, 0, -o], atol=1e-6)
    # 3 px off-centre
    a3 = o * np.exp(-0.5 * 9 / sig2)
    np.testing.assert_allclose(f["opacity"][0, 32, 35], a3, rtol=1e-4)
    # untouched pixels of an active tile: T clamps to 1-1e-6 (forward.cu:484)
    np.testing.assert_allclose(f["opacity"][0, 47, 47], 1e-6, rtol=0.1)
    assert f["n_contrib"].reshape(H, W)[32, 32] == 1


def test_tilted_surfel_depth_plane():
    """A tilted planar surfel: the plane-corrected depth image must follow the analytic plane through the pixel
    rays (that is what the local homography Jinv is for), to first order in the pixel offset."""
    W = H = 96
    cam = syn.Camera(W, H, 80.0, 80.0, 48.0, 48.0)
    z0 = 2.0
    ang = np.deg2rad(35.0)
    n = np.array([np.sin(ang), 0, -np.cos(ang)])      # tilted about y, still facing the camera
    q = syn._quat_z_to(n[None, :])
    xyz = np.array([[0, 0, z0]], np.float32)
    scales = np.array([[0.25, 0.25, 0]], np.float32)
    sh = np.zeros((1, 1, 3), np.float32)
    oc = orc.cam_from_synthetic(cam, 0, 1)
    f = orc.forward(oc, xyz, scales, q, np.array([[0.95]], np.float32), sh)
    assert f["radii"][0] > 0
    for (px, py) in ((48, 48), (52, 48), (44, 50), (50, 44)):
        ray = np.array([(px - cam.cx) / cam.fx, (py - cam.cy) / cam.fy, 1.0])
        t = (n @ np.array([0, 0, z0])) / (n @ ray)       # ray-plane intersection depth
        got = f["depth"][0, py, px]
        assert f["opacity"][0, py, px] > 0.05
        assert abs(got - t) <= 2e-3 * t, (px, py, got, t)


# ------------------------------------------------------------------------------------ structure
def test_binning_is_sorted_and_consistent():
    cam, sc, g, bg, mask, deg, f, _ = util.oracle_run("c1_posed_bg", backward=False)
    I = f["num_rendered"]
    assert I == int(f["tiles_touched"].sum()) == len(f["point_list"])
    keys = f["point_list_keys"]
    assert np.all(keys[1:] >= keys[:-1])
    depth_bits = bits(f["depths"])[f["point_list"]]
    assert np.array_equal((keys & 0xFFFFFFFF).astype(np.uint32), depth_bits)
    tiles = (keys >> 32).astype(np.int64)
    same = tiles[1:] == tiles[:-1]
    tie = same & (keys[1:] == keys[:-1])
    assert np.all(f["point_list"][1:][tie] > f["point_list"][:-1][tie])       # stable: ties keep surfel-id order
    for t in range(f["ranges"].shape[0]):
        a, b = f["ranges"][t]
        assert np.all(tiles[a:b] == t)
    nonempty = np.nonzero(f["ranges"][:, 0] != f["ranges"][:, 1])[0]
    assert np.array_equal(f["tile_indices"][:f["tile_num"]], nonempty)
    assert np.all(f["tile_indices"][f["tile_num"]:] == -1)


def test_tile_mask_restricts_instances_and_is_linear():
    name = "small_deg1"
    cam, sc, g, bg, _, deg = util.case_inputs(name)
    oc = orc.cam_from_synthetic(cam, deg, sc["shs"].shape[1], bg=bg)
    args = (sc["xyz"], sc["scales"], sc["rotations"], sc["opacity"], sc["shs"])
    full = orc.forward(oc, *args)
    ty, tx = cam.tiles
    parts = []
    for r in range(2):
        m = np.zeros((ty, tx), np.int32)
        m[r::2] = 1
        parts.append(orc.forward(oc, *args, tile_mask=m))
    assert parts[0]["num_rendered"] + parts[1]["num_rendered"] == full["num_rendered"]
    assert np.array_equal(parts[0]["color"] + parts[1]["color"], full["color"])
    assert np.array_equal(parts[0]["radii"], full["radii"])     # radii do not depend on the mask
    gb = [orc.backward(oc, p, sc["xyz"], sc["scales"], sc["rotations"], sc["shs"], g["color"], g["normal"],
                       g["depth"], g["opacity"]) for p in parts + [full]]
    for k in ("dL_dmean2D", "dL_dconic", "dL_dcolors", "dL_dopacity"):
        assert rel_err(gb[0][k] + gb[1][k], gb[2][k]) <= 1e-6, k


def test_empty_and_culled_inputs():
    cam = syn.default_camera(64, 48)
    oc = orc.cam_from_synthetic(cam, 0, 1)
    z = lambda *s: np.zeros(s, np.float32)
    f = orc.forward(oc, z(0, 3), z(0, 3), z(0, 4), z(0, 1), z(0, 1, 3))
    assert f["num_rendered"] == 0 and f["tile_num"] == 0 and not f["color"].any()
    sc = syn.make_scene(500, cam, layers=1, sh_degree=0)
    xyz = sc["xyz"].copy()
    xyz[:, 2] *= -1                                         # behind the camera
    f = orc.forward(oc, xyz, sc["scales"], sc["rotations"], sc["opacity"], sc["shs"])
    assert f["num_rendered"] == 0 and not f["active_mask"].any() and not f["radii"].any()
    rot = sc["rotations"].copy()
    rot[:, 1:] *= -1                                        # conjugate: normals flip away from the camera
    f = orc.forward(oc, sc["xyz"], sc["scales"], rot, sc["opacity"], sc["shs"])
    assert f["active_mask"].sum() > 0 and f["radii"].sum() >= 0


def test_oracle_gradient_of_true_paths_by_finite_differences():
    """dL/dSH and dL/dopacity are true derivatives in the reference (unlike normal/depth->mean2D, SURVEY 8 a-bis):
    central differences of the oracle's own forward must reproduce the oracle's backward for them."""
    W, H, P = 48, 32, 60
    cam = syn.default_camera(W, H)
    sc = syn.make_scene(P, cam, layers=2, sh_degree=1, seed=11)
    sc["opacity"] = np.clip(sc["opacity"], 0.3, 0.9).astype(np.float32)      # keep alpha below the 0.99 clamp
    g = syn.make_pixel_grads(cam, seed=12, with_opacity=True)
    oc = orc.cam_from_synthetic(cam, 1, 4)

    def loss(shs, opac):
        f = orc.forward(oc, sc["xyz"], sc["scales"], sc["rotations"], opac, shs)
        return float((f["color"].astype(np.float64) * g["color"]).sum() + (f["opacity"].astype(np.float64) * g["opacity"]).sum()
                     + (f["out_normal"].astype(np.float64) * g["normal"]).sum()
                     + (f["depth"].astype(np.float64) * g["depth"]).sum()), f

    _, f0 = loss(sc["shs"], sc["opacity"])
    b = orc.backward(oc, f0, sc["xyz"], sc["scales"], sc["rotations"], sc["shs"], g["color"], g["normal"], g["depth"],
                     g["opacity"])
    vis = np.nonzero(f0["radii"] > 0)[0]
    rng = np.random.default_rng(0)
    checked = 0
    for i in rng.choice(vis, size=min(8, len(vis)), replace=False):
        # SH: the forward is linear in the coefficients (away from the clamp) -> large step, exact difference
        if not f0["clamped"][i].any():
            k, ch = int(rng.integers(0, 4)), int(rng.integers(0, 3))
            h = 0.25
            sp, sm = sc["shs"].copy(), sc["shs"].copy()
            sp[i, k, ch] += h
            sm[i, k, ch] -= h
            fd = (loss(sp, sc["opacity"])[0] - loss(sm, sc["opacity"])[0]) / (2 * h)
            an = float(b["dL_dsh"][i, k, ch])
            assert abs(fd - an) <= 2e-3 * max(abs(an), np.abs(b["dL_dsh"]).max() * 1e-2), ("sh", i, fd, an)
            checked += 1
        h = 2e-3
        op, om = sc["opacity"].copy(), sc["opacity"].copy()
        op[i] += h
        om[i] -= h
        fd = (loss(sc["shs"], op)[0] - loss(sc["shs"], om)[0]) / (2 * h)
        an = float(b["dL_dopacity"][i, 0])
        assert abs(fd - an) <= 5e-2 * max(abs(an), np.abs(b["dL_dopacity"]).max() * 5e-2), ("opacity", i, fd, an)
        checked += 1
    assert checked >= 8


# ------------------------------------------------------------------------------------ product math on the host
def _emu_forward(emu, cam, sc, deg, bg):
    P, M = sc["xyz"].shape[0], sc["shs"].shape[1]
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    F = C.c_float
    radii = np.zeros(P, np.int32)
    act = np.zeros(P, np.uint8)
    rec = np.zeros((P, 16), np.float32)
    c3 = np.zeros((P, 6), np.float32)
    cl = np.zeros(P, np.uint8)
    rect = np.zeros((P, 4), np.int32)
    view, proj, cp = cam.viewmatrix.reshape(-1).copy(), cam.projmatrix.reshape(-1).copy(), cam.campos.copy()
    bg = np.ascontiguousarray(bg, np.float32)
    op = sc["opacity"].reshape(-1).copy()
    emu.emu_surfel_forward(P, cam.width, cam.height, deg, M, F(cam.tanfovx), F(cam.tanfovy), F(cam.cx), F(cam.cy),
                           F(1.0), p(view), p(proj), p(cp), p(bg), p(sc["xyz"]), p(sc["scales"]), p(sc["rotations"]),
                           p(op), p(sc["shs"]), None, p(radii), p(act), p(rec), p(c3), p(cl), p(rect))
    return radii, act, rec, c3, cl, rect


@pytest.mark.parametrize("variant", ["scene", "big_splats", "unnormalised_quaternions", "near_and_behind", "posed"])
def test_footprint_bound_contains_the_exact_tile_rectangle(hostemu, variant):
    """surfel_bound_rect (the sharded projection's pre-cull, egs_surfel_math.cuh) on the CPU: wherever it decides, its
    tile rectangle contains the exact rectangle the product's surfel_forward computes -- so a surfel it rules out for a
    rank can never have produced an instance there.  Also: it is not vacuous (it is much smaller than the grid)."""
    from eggfusion_b200 import synthetic as syn
    W, H, P = 640, 368, 20000
    base = syn.default_camera(W, H)
    cam = base if variant != "posed" else syn.default_camera(W, H, syn.look_from((0.3, -0.2, 0.25), 0.2, -0.15))
    sc = syn.make_scene(P, base, layers=3, sh_degree=0)
    rng = np.random.default_rng(11)
    if variant == "big_splats":
        sc["scales"][:, :2] *= rng.uniform(1.0, 40.0, size=(P, 1)).astype(np.float32)
    if variant == "unnormalised_quaternions":
        sc["rotations"] *= rng.uniform(0.3, 2.5, size=(P, 1)).astype(np.float32)
    if variant == "near_and_behind":
        sc["xyz"][:, 2] = rng.uniform(-0.3, 1.2, size=P).astype(np.float32)      # across the near plane and behind
        sc["xyz"][:, :2] *= 0.4                                                    # so that many still land on screen
    radii, act, rec, c3, cl, rect = _emu_forward(hostemu, cam, sc, 0, np.zeros(3, np.float32))
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    F = C.c_float
    brect = np.zeros((P, 4), np.int32)
    decided = np.zeros(P, np.uint8)
    view, proj, cp = cam.viewmatrix.reshape(-1).copy(), cam.projmatrix.reshape(-1).copy(), cam.campos.copy()
    hostemu.emu_surfel_bound_rect(P, W, H, F(cam.tanfovx), F(cam.tanfovy), F(cam.cx), F(cam.cy), F(1.0), p(view), p(proj),
                                  p(cp), p(sc["xyz"]), p(sc["scales"]), p(sc["rotations"]), p(brect), p(decided))
    vis = radii > 0
    area = (rect[:, 2] - rect[:, 0]) * (rect[:, 3] - rect[:, 1])
    check = vis & (area > 0) & (decided != 0)
    assert int(check.sum()) > (100 if variant == "near_and_behind" else 1000)
    e, b_ = rect[check], brect[check]
    inside = (b_[:, 0] <= e[:, 0]) & (b_[:, 1] <= e[:, 1]) & (b_[:, 2] >= e[:, 2]) & (b_[:, 3] >= e[:, 3])
    assert bool(inside.all()), (variant, int((~inside).sum()), e[~inside][:3], b_[~inside][:3])
    # every visible surfel with a non-empty rectangle the bound does NOT decide is kept by the caller: nothing to check;
    # but the bound must decide almost all of them and must not be the whole grid
    assert float((decided[vis & (area > 0)] != 0).mean()) > 0.95
    if variant in ("scene", "posed"):
        barea = (b_[:, 2] - b_[:, 0]) * (b_[:, 3] - b_[:, 1])
        gx, gy = (W + 15) // 16, (H + 15) // 16
        assert float(np.median(barea)) <= 4.0 * float(np.median(area[check])) + 4.0
        assert float(barea.mean()) < 0.1 * gx * gy


@pytest.mark.parametrize("name", ["c1_posed_bg", "small_deg0_ragged", "small_deg2"])
def test_product_surfel_math_on_host_matches_oracle(hostemu, name):
    """egs_surfel_math.cuh compiled for the CPU: bit-exact index-critical quantities, tight tolerance elsewhere."""
    cam, sc, g, bg, mask, deg, f, b = util.oracle_run(name)
    radii, act, rec, c3, cl, rect = _emu_forward(hostemu, cam, sc, deg, bg)
    assert np.array_equal(radii, f["radii"]) and np.array_equal(act, f["active_mask"])
    vis = radii > 0
    assert np.array_equal(bits(rec[vis, 0:2]), bits(f["means2D"][vis]))
    assert np.array_equal(bits(rec[vis][:, [4, 5, 6]]), bits(f["conic_opacity"][vis][:, :3]))
    assert np.array_equal(bits(rec[vis, 7]), bits(f["depths"][vis]))
    assert np.array_equal(bits(c3[vis]), bits(f["cov3D"][vis]))
    assert np.array_equal(bits(rec[vis][:, 13:16]), bits(f["normal"][vis]))
    assert rel_err(rec[vis][:, [10, 11, 12]], f["rgb"][vis]) <= 1e-6
    J = f["Jinv"][vis].astype(np.float64)
    assert rel_err(rec[vis, 8], J[:, 0] * J[:, 6] + J[:, 2] * J[:, 9]) <= 1e-5
    assert rel_err(rec[vis, 9], J[:, 1] * J[:, 6] + J[:, 3] * J[:, 9]) <= 1e-5
    # tile rectangles reproduce tiles_touched under an all-ones mask
    area = (rect[:, 2] - rect[:, 0]) * (rect[:, 3] - rect[:, 1])
    if mask.all():
        assert np.array_equal(area[vis].astype(np.uint32), f["tiles_touched"][vis])
    # conservative extent box really bounds the alpha >= 1/255 region: check against the per-pixel rule
    ext = rec[vis, 2].view(np.uint32)
    hx, hy = (ext & 0xFFFF) / 8.0, (ext >> 16) / 8.0
    co = f["conic_opacity"][vis].astype(np.float64)
    tau = 2 * np.log(np.maximum(255 * co[:, 3], 1.0))
    det = co[:, 0] * co[:, 2] - co[:, 1] ** 2
    ex, ey = np.sqrt(tau * co[:, 2] / det), np.sqrt(tau * co[:, 0] / det)    # exact half extents of the ellipse
    assert np.all(hx >= ex) and np.all(hy >= ey)

    # backward of the per-surfel stage
    P, M = sc["xyz"].shape[0], sc["shs"].shape[1]
    sg = util.screen_block_from_oracle(b, P)
    dm, dsh = np.zeros((P, 3), np.float32), np.zeros((P, M, 3), np.float32)
    dsc, dr, dc = np.zeros((P, 3), np.float32), np.zeros((P, 4), np.float32), np.zeros((P, 6), np.float32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    F = C.c_float
    view, proj, cp = cam.viewmatrix.reshape(-1).copy(), cam.projmatrix.reshape(-1).copy(), cam.campos.copy()
    bgc = np.ascontiguousarray(bg, np.float32)
    hostemu.emu_surfel_backward(P, cam.width, cam.height, deg, M, F(cam.tanfovx), F(cam.tanfovy), F(cam.cx), F(cam.cy),
                                F(1.0), p(view), p(proj), p(cp), p(bgc), p(sc["xyz"]), p(sc["scales"]),
                                p(sc["rotations"]), p(sc["shs"]), p(f["radii"]), p(f["cov3D"]), p(cl), p(sg), p(dm),
                                p(dsh), p(dsc), p(dr), p(dc))
    for mine, theirs in ((dm, "dL_dmeans3D"), (dsh, "dL_dsh"), (dsc, "dL_dscales"), (dr, "dL_drotations"),
                         (dc, "dL_dcov3D")):
        assert rel_err(mine, b[theirs]) <= 1e-6, theirs


def test_mark_visible_rule():
    cam = syn.default_camera(64, 48)
    pts = np.array([[0, 0, 1.0], [0, 0, 0.1], [5, 0, 1.0], [0, 0, -1.0], [0.5, 0.3, 1.0]], np.float32)
    vis = orc.mark_visible(pts, cam.viewmatrix, cam.projmatrix)
    assert vis.tolist() == [True, False, False, False, True]


def test_fusion_oracle_matches_reference_golden():
    """project_surfels_to_frame / preprocess_surfels: oracle vs the unmodified reference (B200 golden)."""
    path = util.golden_path("fusion_320x240")
    assert os.path.exists(path)
    G = np.load(path)
    cam, fc = util.fusion_inputs()
    imap, dbuf = orc.project_surfels(fc["points"], fc["rotations"], fc["stable_mask"], fc["intrinsic"], cam.viewmatrix,
                                     cam.projmatrix, cam.height, cam.width)
    assert np.array_equal(dbuf.view(np.uint32), G["depth_buffer"].view(np.uint32))
    bad = imap != G["index_map"]
    assert bad.mean() <= 0.01
    z = fc["points"] @ cam.viewmatrix[:3, 2] + cam.viewmatrix[3, 2]
    assert (z[G["index_map"][bad]] > dbuf[bad]).all()        # the reference's differing ids are stale (its race)
    p, r, s2, inv, surf = orc.fuse_surfels(fc["points"], fc["rotations"], fc["sigma2"], fc["intrinsic"],
                                           cam.viewmatrix, cam.projmatrix, fc["frame_vmap"], fc["frame_nmap"],
                                           fc["frame_dmap"], fc["frame_mask"], G["index_map"],
                                           fc["fusion_dist_thres"], fc["alpha_p"], fc["alpha_n"])
    assert np.array_equal(inv, G["inview_mask"]) and np.array_equal(surf, G["surface_mask"])
    moved = np.abs(p - fc["points"]).max(1) > 0
    assert np.array_equal(moved, np.abs(G["points"] - fc["points"]).max(1) > 0)
    rotated = np.abs(r - fc["rotations"]).max(1) > 0
    assert np.array_equal(rotated, np.abs(G["rotations"] - fc["rotations"]).max(1) > 0)
    assert moved.mean() > 0.3 and rotated.mean() > 0.3 and surf.mean() > 0.1
    assert rel_err(p, G["points"]) <= 1e-6 and rel_err(s2, G["sigma2"]) <= 1e-6
    assert np.abs(r - G["rotations"]).max() <= 2e-5


def test_tracking_oracle_matches_reference_golden():
    """Dense-tracking image utilities: oracle vs the reference's own kernels (tracking.cu compiled with a stub
    Eigen header, run on B200).  solve_block has no reference pin (Eigen is absent): it is checked against numpy."""
    path = util.golden_path("tracking_161x119")
    assert os.path.exists(path)
    G = np.load(path)
    ti = util.tracking_inputs()
    fx, fy, cx, cy = ti["intr"]
    ref = {"bilateral": orc.bilateral_filter(ti["depth"], 13, 0.03, 4.5), "gaussian": orc.gaussian_filter(ti["rgb"], 5, 1.5),
           "down1": orc.gaussian_downsample(ti["gray"][..., None]), "down3": orc.gaussian_downsample(ti["rgb"])}
    ref["grad_x"], ref["grad_y"] = orc.compute_gradient(ti["gray"])
    ref["vertex"], ref["normal"] = orc.compute_vertex_and_normal(ti["depth"], fx, fy, cx, cy)
    for k in ("bilateral", "gaussian", "down1", "down3", "grad_x", "grad_y"):
        assert rel_err(ref[k], G[k]) <= 1e-6, k
    assert np.array_equal(ref["vertex"].view(np.uint32), G["vertex"].view(np.uint32))
    assert np.array_equal(np.all(ref["normal"] == 0, -1), np.all(G["normal"] == 0, -1))
    assert np.abs(ref["normal"] - G["normal"]).max() <= 1e-6
    A = np.diag(np.arange(1, 7)).astype(np.float32)
    np.testing.assert_allclose(orc.solve_block(A, np.ones(6, np.float32), 0.0), 1.0 / np.arange(1, 7), rtol=1e-6)
