"""Host-side logic that needs no GPU: mesh IO, ICP pre-processing, sharding (gloo, world 2),
configuration mirror, and the no-CPU-fallback guarantees."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from followmyhold_b200 import meshio
from followmyhold_b200.alignment import mesh_align as MA
from followmyhold_b200.guidance.config import OptimizationConfig
from followmyhold_b200.synthetic import icosphere, standin_hand_mesh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ply_obj_roundtrip(tmp_path):
    v, f = standin_hand_mesh()
    p = str(tmp_path / "m.ply")
    meshio.write_ply(p, v, f)
    g = meshio.load(p)
    assert isinstance(g, meshio.TriMesh) and np.array_equal(g.faces, f) and np.allclose(g.vertices, v)
    pc = str(tmp_path / "c.ply")
    meshio.write_ply(pc, v)
    g = meshio.load(pc)
    assert isinstance(g, meshio.PointCloud) and g.vertices.shape == (778, 3)
    o = str(tmp_path / "m.obj")
    meshio.write_obj(o, v, f)
    g = meshio.load(o)
    assert np.array_equal(g.faces, f) and np.allclose(g.vertices, v, atol=1e-6)
    # ascii PLY with extra vertex properties (MoGe pointcloud.ply carries colours + normals)
    a = str(tmp_path / "a.ply")
    with open(a, "w") as fh:
        fh.write("ply\nformat ascii 1.0\ncomment x\nelement vertex 3\nproperty float x\nproperty float y\nproperty float z\n"
                 "property uchar red\nproperty uchar green\nproperty uchar blue\nelement face 1\n"
                 "property list uchar int vertex_indices\nend_header\n0 0 0 1 2 3\n1 0 0 4 5 6\n0 1 0 7 8 9\n3 0 1 2\n")
    g = meshio.load(a)
    assert g.vertices.shape == (3, 3) and g.faces.tolist() == [[0, 1, 2]]
    # binary PLY with extra vertex properties, no faces
    b = str(tmp_path / "b.ply")
    rec = np.zeros(5, dtype=[("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("r", "u1"), ("nx", "<f4")])
    rec["x"] = np.arange(5); rec["z"] = 2
    with open(b, "wb") as fh:
        fh.write(b"ply\nformat binary_little_endian 1.0\nelement vertex 5\nproperty float x\nproperty float y\n"
                 b"property float z\nproperty uchar r\nproperty float nx\nend_header\n")
        fh.write(rec.tobytes())
    g = meshio.load(b)
    assert isinstance(g, meshio.PointCloud) and np.allclose(g.vertices[:, 0], np.arange(5)) and np.allclose(g.vertices[:, 2], 2)


def test_centroid_scale_and_init_transform():
    v, f = icosphere(2, 2.0)
    m = meshio.TriMesh(v.astype(np.float64) + np.array([1.0, 2.0, 3.0]), f.astype(np.int64))
    c, s = MA.get_centroid_scale(m)
    assert np.allclose(c, [1, 2, 3], atol=1e-6) and abs(s - np.linalg.norm(m.vertices.max(0) - m.vertices.min(0))) < 1e-12
    pc = meshio.PointCloud(m.vertices * 0.5)
    c2, s2 = MA.get_centroid_scale(pc)
    assert np.allclose(c2, pc.vertices.mean(0))
    T = MA.compute_init_transform(m, pc, fixed_scale=False)
    out = meshio.transform_points(m.vertices, T)
    # centroid lands on the target centroid and the bbox diagonal matches
    assert np.allclose(meshio.TriMesh(out, m.faces).centroid, c2, atol=1e-9)
    assert abs(np.linalg.norm(out.max(0) - out.min(0)) - s2) < 1e-9
    Tf = MA.compute_init_transform(m, pc, fixed_scale=True)
    assert np.allclose(Tf[:3, :3], np.eye(3)) and np.allclose(Tf[:3, 3], c2 - c)
    assert len(MA.get_all_axis_aligned_rotations()) == 9 and len(MA.get_all_axis_aligned_reflections()) == 7


def test_sample_surface_even_properties():
    v, f = icosphere(3, 1.0)
    m = meshio.TriMesh(v.astype(np.float64), f.astype(np.int64))
    rng = np.random.default_rng(0)
    pts, fi = MA.sample_surface_even(m, 1000, rng)
    assert 0 < len(pts) <= 1000 and len(fi) == len(pts)
    # on the surface of the (faceted) sphere
    assert np.all(np.abs(np.linalg.norm(pts, axis=1) - 1.0) < 0.02)
    # thinning: no two samples closer than the radius sqrt(area / (3 count))
    from scipy.spatial import cKDTree
    radius = np.sqrt(m.area / 3000)
    assert len(cKDTree(pts).query_pairs(radius * 0.999)) == 0
    # seeded -> reproducible
    pts2, _ = MA.sample_surface_even(m, 1000, np.random.default_rng(0))
    assert np.array_equal(pts, pts2)


def test_optimization_config_mirrors_reference_constants():
    c = OptimizationConfig()()
    assert (c.optimization_steps_hand, c.optimization_steps_joint, c.optimization_steps_scale) == (200, 50, 100)
    assert (c.num_inference_steps, c.guidance_start_step, c.handopt_start_step, c.guidance_end_step) == (20, 10, 9, 20)
    assert c.phase2_hand_lrs == {"scale": 1e-4, "trans": 1e-4, "rot": 1e-2}
    assert c.obj_lrs == {"scale": 5e-2, "trans": 1e-2, "rot": 1e-2} and c.noise_obj_lr2 == 1e-2 and c.noise_obj_lr1 == 1e-4
    c50 = OptimizationConfig().with_steps(50)
    assert (c50.guidance_start_step, c50.handopt_start_step) == (25, 24)


def test_product_never_imports_the_oracle_and_has_no_cpu_fallback():
    pkg = os.path.join(ROOT, "followmyhold_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports the oracle"
    from followmyhold_b200 import _lib
    from followmyhold_b200.guidance.engine import GuidanceEngine
    with pytest.raises(_lib.FohoLibraryError):
        GuidanceEngine(1, 32, 778, 1538, 0, device="cpu")
    with pytest.raises(_lib.FohoLibraryError):
        MA.icp_points(np.zeros((4, 3)), np.zeros((4, 3)), 1, 0, device="cpu")


WORKER = r'''
import os, sys, json
sys.path.insert(0, sys.argv[1])
import torch.distributed as dist
from followmyhold_b200.parallel import shard_images, gather_timings, aggregate_throughput
dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{sys.argv[2]}", rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
imgs = [f"{i}_cropped_obj.png" for i in range(7)]
mine = shard_images(imgs, r, 2)
per = gather_timings({"units": float(len(mine)), "seconds": 1.0 + r})
if r == 0:
    print(json.dumps({"mine": mine, "per": per, "agg": aggregate_throughput(per)}))
dist.barrier(); dist.destroy_process_group()
'''


def test_sharding_and_timing_gather_world2_gloo(tmp_path):
    from followmyhold_b200.parallel import shard_images
    imgs = [f"{i}_cropped_obj.png" for i in range(7)]
    parts = [shard_images(imgs, r, 2) for r in range(2)]
    assert sorted(parts[0] + parts[1]) == sorted(imgs) and not set(parts[0]) & set(parts[1])
    script = tmp_path / "w.py"
    script.write_text(WORKER)
    import socket
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(port), str(r)], stdout=subprocess.PIPE, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=120)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs)
    import json
    res = json.loads(outs[0].strip().splitlines()[-1])
    assert res["mine"] == parts[0]
    assert [p["units"] for p in res["per"]] == [4.0, 3.0]
    assert abs(res["agg"] - 7.0 / 2.0) < 1e-12      # all units / slowest rank


def test_cpulist_parser_and_numa_binding_is_best_effort(tmp_path):
    from followmyhold_b200.parallel import bind_to_gpu_numa, parse_cpulist
    assert parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parse_cpulist("") == []
    # no GPU here: the helper must not raise and must report that nothing was bound
    info = bind_to_gpu_numa(0, sysfs=str(tmp_path))
    assert info["bound"] is False


def test_delaunay_graph_makes_greedy_descent_exact():
    """The claim k_chamfer_c2h_walk rests on: on the Delaunay neighbour graph of the rest hand, moving to
    the neighbour closest to the query until none is closer ends at the true nearest vertex, from any
    start, for near and far queries alike.  (Emulated here with numpy on the graph the host builds.)"""
    import torch
    from followmyhold_b200.guidance.engine import delaunay_neighbours
    from followmyhold_b200.synthetic import make_guidance_sample
    s = make_guidance_sample(16, 64, 5)
    rest = s.hand_rest.numpy().astype(np.float64)
    off, nbr = delaunay_neighbours(torch.stack([s.hand_rest, s.hand_rest * 1.3 + 0.1]))
    assert off.shape == (2, 779) and off.dtype == torch.int32 and nbr.shape[1] % 8 == 0
    indptr, idx = off[0].numpy(), nbr[0].numpy().view(np.uint16)
    assert indptr[0] == 0 and (np.diff(indptr) >= 3).all()
    # symmetric adjacency
    pairs = {(i, int(j)) for i in range(778) for j in idx[indptr[i]:indptr[i + 1]]}
    assert all((j, i) in pairs for i, j in pairs)
    rng = np.random.default_rng(0)
    c, ext = rest.mean(0), np.ptp(rest, axis=0).max()
    queries = np.concatenate([rest[rng.integers(0, 778, 300)] + 0.02 * ext * rng.normal(size=(300, 3)),
                              c + 0.6 * ext * rng.normal(size=(300, 3)), c + 6.0 * ext * rng.normal(size=(300, 3))])
    for q in queries:
        d2 = ((rest - q) ** 2).sum(1)
        cur = int(rng.integers(0, 778))
        for _ in range(778):
            nb = idx[indptr[cur]:indptr[cur + 1]].astype(np.int64)
            k = nb[d2[nb].argmin()]
            if d2[k] < d2[cur]:
                cur = int(k)
            else:
                break
        assert d2[cur] == d2.min()
    # degenerate input (all points in a plane): no graph, the caller keeps the box search
    flat = s.hand_rest.clone(); flat[:, 2] = 0.0
    assert delaunay_neighbours(flat[None]) is None


def _write_glb(path, prims, nodes):
    """Minimal binary glTF 2.0 writer for the reader's test: prims = [(verts, indices or None, index dtype)]."""
    import json
    import struct
    blob, views, accs, meshes = b"", [], [], []
    for v, idx, dt in prims:
        v = np.asarray(v, "<f4")
        views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": v.nbytes})
        blob += v.tobytes() + b"\0" * ((-v.nbytes) % 4)
        accs.append({"bufferView": len(views) - 1, "componentType": 5126, "count": len(v), "type": "VEC3"})
        prim = {"attributes": {"POSITION": len(accs) - 1}}
        if idx is not None:
            a = np.asarray(idx, dt).reshape(-1)
            views.append({"buffer": 0, "byteOffset": len(blob), "byteLength": a.nbytes})
            blob += a.tobytes() + b"\0" * ((-a.nbytes) % 4)
            accs.append({"bufferView": len(views) - 1, "componentType": {"u1": 5121, "<u2": 5123, "<u4": 5125}[dt],
                         "count": int(a.size), "type": "SCALAR"})
            prim["indices"] = len(accs) - 1
        else:
            prim["mode"] = 0
        meshes.append({"primitives": [prim]})
    g = {"asset": {"version": "2.0"}, "scene": 0, "scenes": [{"nodes": [0]}], "nodes": nodes, "meshes": meshes,
         "accessors": accs, "bufferViews": views, "buffers": [{"byteLength": len(blob)}]}
    js = json.dumps(g).encode()
    js += b" " * ((-len(js)) % 4)
    with open(path, "wb") as f:
        f.write(b"glTF" + struct.pack("<II", 2, 28 + len(js) + len(blob)) + struct.pack("<II", len(js), 0x4E4F534A) + js
                + struct.pack("<II", len(blob), 0x004E4942) + blob)


def test_glb_reader_geometry_and_node_transforms(tmp_path):
    """``mesh.glb`` (MoGe's output, which the reference loads at guidance/run.py:215 and alignment/h2m.py:27-31):
    positions + indices of every triangle primitive, scene-graph transforms applied, concatenated."""
    rng = np.random.default_rng(0)
    v0, f0 = rng.normal(size=(5, 3)), [[0, 1, 2], [2, 3, 4]]
    v1, f1 = rng.normal(size=(4, 3)), [[0, 1, 2], [1, 2, 3]]
    nodes = [{"children": [1, 2], "translation": [1.0, 2.0, 3.0]}, {"mesh": 0},
             {"mesh": 1, "scale": [2, 2, 2], "rotation": [0, 0, 0.7071067811865476, 0.7071067811865476]}]   # 90 deg about z
    _write_glb(str(tmp_path / "mesh.glb"), [(v0, f0, "<u2"), (v1, f1, "<u4")], nodes)
    m = meshio.load(str(tmp_path / "mesh.glb"))
    assert isinstance(m, meshio.TriMesh) and m.vertices.shape == (9, 3) and m.faces.shape == (4, 3)
    Rz = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    want = np.concatenate([v0.astype("f4") + [1, 2, 3], (2 * v1.astype("f4")) @ Rz.T + [1, 2, 3]])
    got_tris = np.sort(np.round(m.vertices[m.faces].reshape(4, -1), 5), axis=0)
    want_tris = np.sort(np.round(np.concatenate([want[:5][np.array(f0)], want[5:][np.array(f1)]]).reshape(4, -1), 5), axis=0)
    assert np.allclose(got_tris, want_tris, atol=1e-5)
    # points only -> PointCloud; byte indices; garbage -> ValueError
    _write_glb(str(tmp_path / "p.glb"), [(v0, None, None)], [{"mesh": 0}])
    pc = meshio.load(str(tmp_path / "p.glb"))
    assert isinstance(pc, meshio.PointCloud) and np.allclose(pc.vertices, v0.astype("f4"), atol=1e-6)
    _write_glb(str(tmp_path / "b.glb"), [(v1, f1, "u1")], [{"mesh": 0}])
    assert meshio.load(str(tmp_path / "b.glb")).faces.tolist() == f1
    (tmp_path / "bad.glb").write_bytes(b"not a gltf file at all.....")
    with pytest.raises(ValueError):
        meshio.load(str(tmp_path / "bad.glb"))
    # the guidance stage prefers mesh.glb, like the reference
    from followmyhold_b200.guidance import run as R
    src = inspect_source = __import__("inspect").getsource(R.load_image_inputs)
    assert src.index('"mesh.glb"') < src.index('"pointcloud.ply"')


def test_align_meshes_many_host_logic_equals_one_at_a_time(tmp_path, monkeypatch):
    """Grouping, seeds, composition and file writing of the batched alignment, with the CPU oracle standing in
    for the device loop (tests only): every image gets what one ``align_meshes_impl`` call gives it."""
    from followmyhold_b200 import meshio
    from followmyhold_b200.synthetic import icosphere, standin_hand_mesh
    from oracle import icp_oracle as IO

    def one(src, tgt, n_iter, n_out, fixed_scale=False, min_scale=0.5, max_scale=2.0, device=None, return_history=False):
        return IO.icp_points(src, tgt, n_iter, n_out, fixed_scale, min_scale, max_scale)

    def many(problems, n_iter, n_outliers, fixed_scale=False, min_scale=0.5, max_scale=2.0, device=None):
        outs = [n_outliers] * len(problems) if isinstance(n_outliers, int) else n_outliers
        return [one(s, t, n_iter, o, fixed_scale, min_scale, max_scale) for (s, t), o in zip(problems, outs)]

    monkeypatch.setattr(MA, "icp_points", one)
    monkeypatch.setattr(MA, "icp_points_many", many)
    # the device thinning returns the host statement's mask bit for bit (test_gpu_icp.py); no GPU here
    monkeypatch.setattr(MA, "remove_close_device", lambda pts, radius, device=None: MA.remove_close(pts, radius))
    hv, hf = standin_hand_mesh(0.35)
    jobs = []
    for k in range(3):
        v, f = icosphere(2, 0.4)
        src = tmp_path / f"{k}_src.obj"; tgt = tmp_path / f"{k}_tgt.ply"
        meshio.write_obj(str(src), hv * (1.1 + 0.1 * k) + 0.05 * k, hf)
        meshio.write_ply(str(tgt), v.astype(np.float64) * np.array([1.0, 0.8, 0.6 + 0.1 * k]), f)
        jobs.append((str(src), str(tgt), str(tmp_path / f"many_{k}"), str(tmp_path / f"many_{k}.ply")))
    kw = dict(fixed_scale=False, outliers=0.2, test_rotations=False, test_reflections=True, on_surface=False,
              iterations_coarse=4, count_source_coarse=150, count_target_coarse=300, iterations_fine=5,
              count_source_fine=200, count_target_fine=400, min_scale=0.7, max_scale=3.0, plot=False, seed=3)
    finals = MA.align_meshes_many(jobs, concurrent=2, workers=3, **kw)
    serial = MA.align_meshes_many([(a, b, None, None) for a, b, _, _ in jobs], concurrent=3, workers=1, **kw)
    assert all(np.array_equal(x, y) for x, y in zip(finals, serial))       # host threads do not change results
    assert 1 <= MA.host_workers(5) <= 5 and MA.host_workers(0) == 1
    for k, (src, tgt, _, _) in enumerate(jobs):
        ref = MA.align_meshes_impl(src, tgt, str(tmp_path / f"one_{k}"), str(tmp_path / f"one_{k}.ply"), **kw)
        assert np.array_equal(finals[k], ref)
        assert np.array_equal(np.load(tmp_path / f"many_{k}.npy"), np.load(tmp_path / f"one_{k}.npy"))
        assert np.array_equal(meshio.load(str(tmp_path / f"many_{k}.ply")).vertices,
                              meshio.load(str(tmp_path / f"one_{k}.ply")).vertices)
    assert MA.align_meshes_many([], **kw) == []


def test_alignment_stage_glue_on_files_with_the_oracle_loop(tmp_path, monkeypatch, capsys):
    """``h2m.run`` / ``mano.run`` end to end on files (target lookup order, file names, the callers' ICP constants,
    rank sharding) with the CPU oracle standing in for the device loop: same artefacts as one reference-style
    ``align_meshes_impl`` call per image with the constants of h2m.py:35-54 / mano.py:24-43."""
    from followmyhold_b200 import meshio
    from followmyhold_b200.alignment import h2m, mano
    from followmyhold_b200.synthetic import icosphere, standin_hand_mesh
    from oracle import icp_oracle as IO

    def one(src, tgt, n_iter, n_out, fixed_scale=False, min_scale=0.5, max_scale=2.0, device=None, return_history=False):
        return IO.icp_points(src, tgt, n_iter, n_out, fixed_scale, min_scale, max_scale)

    monkeypatch.setattr(MA, "remove_close_device", lambda pts, radius, device=None: MA.remove_close(pts, radius))
    monkeypatch.setattr(MA, "icp_points", one)
    monkeypatch.setattr(MA, "icp_points_many", lambda probs, n_iter, n_out, fs=False, lo=0.5, hi=2.0, device=None:
                        [one(s, t, n_iter, o, fs, lo, hi) for (s, t), o in zip(probs, n_out)])
    assert MA.STAGE_ICP_KWARGS == dict(fixed_scale=False, outliers=0.2, test_rotations=False, test_reflections=False,
                                       on_surface=False, iterations_coarse=50, count_source_coarse=1000,
                                       count_target_coarse=5000, iterations_fine=100, count_source_fine=5000,
                                       count_target_fine=10000, min_scale=0.7, max_scale=3.0, plot=False)
    hv, hf = standin_hand_mesh(0.35)
    hun = tmp_path / "hun"; ham = tmp_path / "hamer"; hun.mkdir(); ham.mkdir()
    rng = np.random.default_rng(0)
    for k, i in enumerate(("04", "09")):
        v, f = icosphere(2, 0.4)
        v = v.astype(np.float64) * np.array([1.0, 0.75, 0.5 + 0.1 * k])
        meshio.write_ply(str(hun / f"{i}_hoi_mesh.ply"), v, f)
        md = tmp_path / "moge" / f"{i}_cropped_hoi"; md.mkdir(parents=True)
        cloud = v[f][rng.integers(0, len(f), 3000)].mean(1) * 0.4 + np.array([0.1, 0.0, -1.2])
        # image 04 has both candidates: pointcloud.ply must win over mesh.glb; 09 has only a decoy name
        meshio.write_ply(str(md / "pointcloud.ply"), cloud)
        if k == 0:
            (md / "mesh.glb").write_bytes(b"not read")
        meshio.write_obj(str(ham / f"{i}_hamer.obj"), hv * 1.2 + 0.1, hf)
    (tmp_path / "moge" / "09_cropped_hoi" / "pointcloud.ply").rename(tmp_path / "moge" / "09_cropped_hoi" / "mesh.ply")
    h2m.run(str(hun), str(tmp_path / "moge"), str(tmp_path / "rt"), concurrent=2)
    mano.run(str(ham), str(hun), str(tmp_path / "aligned"), concurrent=1)
    assert sorted(os.listdir(tmp_path / "rt")) == ["04_hoi_mesh.npy", "09_hoi_mesh.npy"]
    assert sorted(os.listdir(tmp_path / "aligned")) == ["04_hamer_aligned_mano.ply", "09_hamer_aligned_mano.ply"]
    for i, tgt in (("04", "pointcloud.ply"), ("09", "mesh.ply")):
        ref = MA.align_meshes_impl(str(hun / f"{i}_hoi_mesh.ply"), str(tmp_path / "moge" / f"{i}_cropped_hoi" / tgt),
                                   None, None, **MA.STAGE_ICP_KWARGS, seed=0)
        got = np.load(tmp_path / "rt" / f"{i}_hoi_mesh.npy")
        assert got.dtype == np.float64 and np.array_equal(got, ref)
        assert 0.3 < np.linalg.norm(got[:3, 0]) < 0.5                                 # the 0.4x similarity is recovered
    # rank 1 of 2 aligns only its share; a frame without MoGe geometry is reported and skipped
    monkeypatch.setenv("RANK", "1"); monkeypatch.setenv("WORLD_SIZE", "2"); monkeypatch.setenv("FOHO_B200_SHARD", "1")
    h2m.run(str(hun), str(tmp_path / "moge"), str(tmp_path / "rt1"))
    assert os.listdir(tmp_path / "rt1") == ["09_hoi_mesh.npy"]
    monkeypatch.delenv("RANK"); monkeypatch.delenv("WORLD_SIZE"); monkeypatch.delenv("FOHO_B200_SHARD")
    import shutil
    shutil.rmtree(tmp_path / "moge" / "04_cropped_hoi")
    capsys.readouterr()
    h2m.run(str(hun), str(tmp_path / "moge"), str(tmp_path / "rt2"))
    assert "No MoGe mesh found for 04" in capsys.readouterr().out and os.listdir(tmp_path / "rt2") == ["09_hoi_mesh.npy"]
    h2m.run(str(tmp_path / "empty"), str(tmp_path / "moge"), str(tmp_path / "rt3"))
    assert "No Hunyuan HOI meshes found" in capsys.readouterr().out
