"""GPU parity of the ICP loop (foho_icp_run through the C-ABI) against the CPU oracle
(scipy cKDTree + restated trimesh procrustes, oracle/icp_oracle.py)."""
import numpy as np
import pytest

from followmyhold_b200.synthetic import icosphere, random_similarity, standin_hand_mesh

pytestmark = pytest.mark.gpu


def _clouds(ns, nt, seed, outlier_frac=0.2):
    rng = np.random.default_rng(seed)
    v, f = standin_hand_mesh(1.0)
    tri = v[f].astype(np.float64)

    def sample(n):
        fi = rng.integers(0, len(f), n)
        bc = rng.random((n, 2))
        flip = bc.sum(1) > 1
        bc[flip] = 1 - bc[flip]
        t = tri[fi]
        return t[:, 0] + bc[:, :1] * (t[:, 1] - t[:, 0]) + bc[:, 1:] * (t[:, 2] - t[:, 0])

    tgt = sample(nt)
    src = sample(ns)
    T = random_similarity(seed, (0.8, 1.3), 0.05)
    # small rotation so plain ICP converges: blend towards identity
    T[:3, :3] = 0.15 * T[:3, :3] + 0.85 * np.linalg.norm(T[:3, 0]) * np.eye(3)
    u, s, vh = np.linalg.svd(T[:3, :3]); T[:3, :3] = (u @ vh) * s.mean()
    Tinv = np.linalg.inv(T)
    src = src @ Tinv[:3, :3].T + Tinv[:3, 3]
    n_out = int(outlier_frac * ns)
    src[:n_out] += rng.normal(scale=0.5, size=(n_out, 3))
    return src, tgt, T


@pytest.mark.parametrize("ns,nt,n_iter", [(1000, 5000, 50), (5000, 10000, 100), (777, 1234, 30), (2000, 70000, 20)])
def test_icp_matches_oracle(ns, nt, n_iter):
    from followmyhold_b200.alignment.mesh_align import icp_points
    from oracle import icp_oracle as O
    src, tgt, T_true = _clouds(ns, nt, seed=ns + nt)
    n_out = int(0.2 * ns)
    T, cost, hist = icp_points(src, tgt, n_iter, n_out, False, 0.7, 3.0, return_history=True)
    To, co, ho, _ = O.icp_points(src, tgt, n_iter, n_out, False, 0.7, 3.0, return_history=True)
    # float64 everywhere; only summation order differs
    np.testing.assert_allclose(hist, ho, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(T, To, rtol=0, atol=1e-8)
    assert abs(cost - co) <= 1e-10 * max(1.0, abs(co))


def test_icp_recovers_known_similarity():
    from followmyhold_b200.alignment.mesh_align import icp_points
    src, tgt, T_true = _clouds(3000, 3000, seed=7, outlier_frac=0.0)
    # exact correspondences: target = T_true(source) -> converges to T_true
    tgt = src @ T_true[:3, :3].T + T_true[:3, 3]
    T, cost = icp_points(src, tgt, 60, 0, False, 0.5, 3.0)
    np.testing.assert_allclose(T, T_true, atol=1e-6)
    assert cost < 1e-6


def test_icp_fixed_scale_and_clip():
    from followmyhold_b200.alignment.mesh_align import icp_points
    from oracle import icp_oracle as O
    src, tgt, _ = _clouds(1500, 4000, seed=3)
    for fixed, lo, hi in ((True, 0.5, 2.0), (False, 0.95, 1.05)):
        T, cost, hist = icp_points(src, tgt, 25, 300, fixed, lo, hi, return_history=True)
        To, co, ho, _ = O.icp_points(src, tgt, 25, 300, fixed, lo, hi, return_history=True)
        np.testing.assert_allclose(hist, ho, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(T, To, atol=1e-8)
        if not fixed:
            sc = np.linalg.norm(T[:3, 0])
            assert lo - 1e-12 <= sc <= hi + 1e-12


def test_align_meshes_impl_files(tmp_path):
    """File-level seam: same artefacts as the reference stage (h2m writes <j>.npy)."""
    from followmyhold_b200 import meshio
    from followmyhold_b200.alignment import h2m, mano
    v, f = icosphere(3, 0.4)
    v = v.astype(np.float64) * np.array([1.0, 0.7, 0.5])
    hv, hf = standin_hand_mesh(0.35)
    hun = tmp_path / "hunyuan"; moge = tmp_path / "moge" / "12_cropped_hoi"; rt = tmp_path / "rt"
    ham = tmp_path / "hamer"; out = tmp_path / "aligned"
    for d in (hun, moge, ham):
        d.mkdir(parents=True)
    meshio.write_ply(str(hun / "12_hoi_mesh.ply"), v, f)
    T = random_similarity(5, (0.3, 0.4), 0.1)
    T[:3, :3] = np.linalg.norm(T[:3, 0]) * np.eye(3)
    T[:3, 3] += np.array([0, 0, -1.5])
    rng = np.random.default_rng(0)
    tri = v[f]
    fi = rng.integers(0, len(f), 20000)
    cloud = tri[fi].mean(1) @ T[:3, :3].T + T[:3, 3]
    meshio.write_ply(str(moge / "pointcloud.ply"), cloud)
    h2m.run(str(hun), str(tmp_path / "moge"), str(rt))
    M = np.load(rt / "12_hoi_mesh.npy")
    assert M.shape == (4, 4) and M.dtype == np.float64
    assert abs(np.linalg.norm(M[:3, 0]) - np.linalg.norm(T[:3, 0])) < 0.05
    meshio.write_obj(str(ham / "12_hamer.obj"), hv * 1.3 + 0.2, hf)
    mano.run(str(ham), str(hun), str(out))
    g = meshio.load(str(out / "12_hamer_aligned_mano.ply"))
    assert g.vertices.shape == (778, 3) and g.faces.shape == (1538, 3)


@pytest.mark.gpu
def test_concurrent_loops_equal_single_loops():
    """icp_points_many (one stream per problem) returns exactly what icp_points returns per problem."""
    from followmyhold_b200.alignment.mesh_align import icp_points, icp_points_many
    rng = np.random.default_rng(5)
    probs = []
    for k in range(4):
        tgt = rng.normal(size=(1500 + 100 * k, 3))
        src = (tgt[:400 + 10 * k] - 0.03) / (1.05 + 0.01 * k) + 0.005 * rng.normal(size=(400 + 10 * k, 3))
        probs.append((src, tgt))
    many = icp_points_many(probs, 15, [80, 82, 84, 86], False, 0.7, 3.0)
    for (src, tgt), n_out, (T, c) in zip(probs, [80, 82, 84, 86], many):
        T1, c1 = icp_points(src, tgt, 15, n_out, False, 0.7, 3.0)
        assert np.array_equal(T, T1) and c == c1


@pytest.mark.gpu
def test_one_launch_for_several_problems_equals_single_runs_at_the_stage_sizes():
    """foho_icp_run_batch divides the SMs between the problems; block sums are formed per 32 consecutive points and
    added in index order, so the number of CTAs a problem gets does not change a single bit."""
    from followmyhold_b200.alignment.mesh_align import icp_points, icp_points_many
    probs, outs = [], []
    for k in range(3):
        src, tgt, _ = _clouds(5000 - 37 * k, 10000 + 11 * k, seed=20 + k)
        probs.append((src, tgt))
        outs.append(int(0.2 * src.shape[0]))
    many = icp_points_many(probs, 40, outs, False, 0.7, 3.0)
    for (src, tgt), n_out, (T, c) in zip(probs, outs, many):
        T1, c1 = icp_points(src, tgt, 40, n_out, False, 0.7, 3.0)
        assert np.array_equal(T, T1) and c == c1


@pytest.mark.gpu
def test_device_thinning_mask_equals_the_host_statement():
    """foho_remove_close (grid + sort) against ``remove_close`` (cKDTree.query_pairs + bincount + argmax, the statement of
    trimesh.points.remove_close): identical masks on surface samples at the stage's four sizes, on clouds with exact
    duplicates, with fewer points than the sort's smallest size, with radius 0, and where the box is far wider than
    16k cells of the radius."""
    from followmyhold_b200.alignment import mesh_align as MA
    from followmyhold_b200.meshio import TriMesh
    from followmyhold_b200.synthetic import icosphere
    v, f = icosphere(3, 0.4)
    mesh = TriMesh(v.astype(np.float64) * np.array([1.0, 0.7, 0.45]), f)
    for count in (1000, 5000, 10000):
        rng = np.random.default_rng(count)
        pts, _ = MA.sample_surface(mesh, 3 * count, rng)
        radius = np.sqrt(mesh.area / (3 * count))
        kept_h, mask_h = MA.remove_close(pts, radius)
        kept_d, mask_d = MA.remove_close_device(pts, radius)
        assert np.array_equal(mask_h, mask_d) and np.array_equal(kept_h, kept_d)
        assert 0.2 * len(pts) < mask_d.sum() < len(pts)
        a, _ = MA.sample_surface_even(mesh, count, np.random.default_rng(1))
        b, _ = MA.sample_surface_even(mesh, count, np.random.default_rng(1), device="cuda:0")
        assert np.array_equal(a, b)
    rng = np.random.default_rng(0)
    dup = rng.normal(size=(700, 3))
    dup = np.concatenate([dup, dup[:200], dup[:50] + 1e-3])                          # exact duplicates and near ones
    for radius in (0.0, 1e-3, 0.2, 5.0):
        assert np.array_equal(MA.remove_close(dup, radius)[1], MA.remove_close_device(dup, radius)[1]), radius
    wide = rng.uniform(-1e3, 1e3, size=(4000, 3))
    wide[::7] = wide[1::7][:len(wide[::7])] + 1e-4 * rng.normal(size=(len(wide[::7]), 3))   # pairs ~1e-4 apart in a 2e3 box
    assert np.array_equal(MA.remove_close(wide, 2e-4)[1], MA.remove_close_device(wide, 2e-4)[1])
    assert MA.remove_close_device(np.zeros((0, 3)), 0.1)[1].shape == (0,)


@pytest.mark.gpu
def test_trim_with_many_equal_distances_matches_the_oracle():
    """Distances that tie at the trim threshold (lattice clouds, an exact offset): the lowest indices among the ties are
    kept, like the stable argsort of the reference's loop (mesh_align.py:114-120)."""
    from followmyhold_b200.alignment.mesh_align import icp_points
    from oracle import icp_oracle as O
    g = np.arange(12, dtype=np.float64)
    tgt = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)          # 1728 lattice points
    rng = np.random.default_rng(3)
    src = tgt[rng.permutation(len(tgt))[:1200]] + np.array([0.25, 0.0, 0.0])          # every distance is exactly 0.25
    T, cost, hist = icp_points(src, tgt, 2, 300, True, 0.5, 2.0, return_history=True)
    To, co, ho, _ = O.icp_points(src, tgt, 2, 300, True, 0.5, 2.0, return_history=True)
    np.testing.assert_allclose(hist[:1], ho[:1], rtol=1e-12)
    np.testing.assert_allclose(T, To, atol=1e-9)


def test_batched_sharded_alignment_stage_equals_one_image_at_a_time(tmp_path, monkeypatch):
    """The stages align the images of a batch with their ICP loops in flight together, and a rank only its
    share ``sorted(meshes)[r::world]`` -- transforms and meshes are bit-equal to one ``align_meshes_impl`` call
    per image (the reference's one-at-a-time loop, h2m.py:17-54 / mano.py:17-43)."""
    from followmyhold_b200 import meshio
    from followmyhold_b200.alignment import h2m, mano
    from followmyhold_b200.alignment.mesh_align import align_meshes_impl
    hv, hf = standin_hand_mesh(0.35)
    hun = tmp_path / "hunyuan"; ham = tmp_path / "hamer"
    hun.mkdir(); ham.mkdir()
    rng = np.random.default_rng(1)
    ids = ["03", "07", "11"]
    for k, i in enumerate(ids):
        v, f = icosphere(3, 0.4)
        v = v.astype(np.float64) * np.array([1.0, 0.7 + 0.05 * k, 0.5])
        meshio.write_ply(str(hun / f"{i}_hoi_mesh.ply"), v, f)
        md = tmp_path / "moge" / f"{i}_cropped_hoi"; md.mkdir(parents=True)
        T = random_similarity(5 + k, (0.3, 0.4), 0.1)
        T[:3, :3] = np.linalg.norm(T[:3, 0]) * np.eye(3)
        T[:3, 3] += np.array([0, 0, -1.5])
        tri = v[f]
        cloud = tri[rng.integers(0, len(f), 12000)].mean(1) @ T[:3, :3].T + T[:3, 3]
        meshio.write_ply(str(md / "pointcloud.ply"), cloud)
        meshio.write_obj(str(ham / f"{i}_hamer.obj"), hv * (1.2 + 0.1 * k) + 0.1 * k, hf)
    kw = dict(fixed_scale=False, outliers=0.2, test_rotations=False, test_reflections=False, on_surface=False,
              iterations_coarse=50, count_source_coarse=1000, count_target_coarse=5000, iterations_fine=100,
              count_source_fine=5000, count_target_fine=10000, min_scale=0.7, max_scale=3.0, plot=False, seed=0)
    # batched (all three in one group), and rank 1 of 2 (only image "07")
    h2m.run(str(hun), str(tmp_path / "moge"), str(tmp_path / "rt"), concurrent=3)
    mano.run(str(ham), str(hun), str(tmp_path / "aligned"), concurrent=2)
    monkeypatch.setenv("RANK", "1"); monkeypatch.setenv("WORLD_SIZE", "2"); monkeypatch.setenv("FOHO_B200_SHARD", "1")
    h2m.run(str(hun), str(tmp_path / "moge"), str(tmp_path / "rt_rank1"))
    monkeypatch.delenv("RANK"); monkeypatch.delenv("WORLD_SIZE"); monkeypatch.delenv("FOHO_B200_SHARD")
    assert sorted(p.name for p in (tmp_path / "rt_rank1").iterdir()) == ["07_hoi_mesh.npy"]
    for i in ids:
        ref = align_meshes_impl(str(hun / f"{i}_hoi_mesh.ply"), str(tmp_path / "moge" / f"{i}_cropped_hoi" / "pointcloud.ply"),
                                str(tmp_path / f"one_{i}"), None, **kw)
        assert np.array_equal(np.load(tmp_path / "rt" / f"{i}_hoi_mesh.npy"), ref)
        align_meshes_impl(str(ham / f"{i}_hamer.obj"), str(hun / f"{i}_hoi_mesh.ply"), None, str(tmp_path / f"one_{i}.ply"), **kw)
        a = meshio.load(str(tmp_path / "aligned" / f"{i}_hamer_aligned_mano.ply")).vertices
        b = meshio.load(str(tmp_path / f"one_{i}.ply")).vertices
        assert np.array_equal(a, b)
    assert np.array_equal(np.load(tmp_path / "rt_rank1" / "07_hoi_mesh.npy"), np.load(tmp_path / "rt" / "07_hoi_mesh.npy"))


@pytest.mark.gpu
def test_batch_launch_with_mixed_paths_and_edge_sizes():
    """One ``foho_icp_run_batch`` call over problems that take different paths: a target below 1 024 points (launch pairs,
    brute-force 1-NN), a large one (persistent loop), a single iteration, no trimming, and a source with fewer points than
    one warp's worth of CTAs -- each equal to its own ``icp_points`` run and to the oracle."""
    from followmyhold_b200.alignment.mesh_align import icp_points, icp_points_many
    from oracle import icp_oracle as O
    specs = [(300, 700, 0.2), (2500, 6000, 0.2), (40, 3000, 0.0), (1500, 1024, 0.1)]
    probs, outs = [], []
    for k, (ns, nt, frac) in enumerate(specs):
        src, tgt, _ = _clouds(ns, nt, seed=50 + k)
        probs.append((src, tgt)); outs.append(int(frac * ns))
    for n_iter in (1, 12):
        many = icp_points_many(probs, n_iter, outs, False, 0.7, 3.0)
        for (src, tgt), n_out, (T, c) in zip(probs, outs, many):
            T1, c1 = icp_points(src, tgt, n_iter, n_out, False, 0.7, 3.0)
            assert np.array_equal(T, T1) and c == c1
            To, co = O.icp_points(src, tgt, n_iter, n_out, False, 0.7, 3.0)[:2]
            np.testing.assert_allclose(T, To, rtol=0, atol=1e-8)
            assert abs(c - co) <= 1e-10 * max(1.0, abs(co))
    assert icp_points_many([], 5, 0) == []
