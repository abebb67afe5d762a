"""Guidance stage mirror (followmyhold_b200.guidance.run): the reference's signature, flags, file
contract and skip rules (src/foho/guidance/run.py:178-289).  Host logic on the CPU; the GPU test
drives three synthetic images through the whole schedule."""
import inspect
import json
import os

import numpy as np
import pytest
import torch

from followmyhold_b200.guidance import run as R
from followmyhold_b200.meshio import load, write_ply
from followmyhold_b200.synthetic import make_guidance_sample, standin_j_regressor

REF_ARGS = ["project_root", "cropped_obj_img_dir", "mask_dir", "moge_out_dir", "hunyuan_hoi_mesh_dir", "hamer_out_dir",
            "h2m_rt_dir", "aligned_mano_dir", "guidance_out_dir", "task_list_file"]      # run.py:188-199


def write_dataset(root, n_images, D=32, P=2000, empty_mask_for=None):
    """Files of the stages upstream of guidance, shaped per SURVEY.md section 8b, for synthetic images."""
    import cv2
    d = {k: os.path.join(root, k) for k in ("img", "mask", "moge", "hun", "hamer", "h2m", "mano", "out")}
    for v in d.values():
        os.makedirs(v, exist_ok=True)
    for k in range(n_images):
        i = f"{k:03d}"
        s = make_guidance_sample(D, P, 200 + k)
        open(os.path.join(d["img"], f"{i}_cropped_obj_1.png"), "wb").close()
        m = np.zeros((64, 64), np.uint8); m[20:40, 20:40] = 255
        cv2.imwrite(os.path.join(d["mask"], f"{i}_cropped_hand_mask.png"), m * 0 if empty_mask_for == k else m)
        cv2.imwrite(os.path.join(d["mask"], f"{i}_cropped_obj_mask.png"), m)
        md = os.path.join(d["moge"], f"{i}_cropped_hoi"); os.makedirs(md, exist_ok=True)
        json.dump({"fov_x": 41.0}, open(os.path.join(md, "fov.json"), "w"))
        write_ply(os.path.join(md, "pointcloud.ply"), s.cloud.numpy().astype(np.float64))
        T = s.T_h2m.numpy().astype(np.float64)
        np.save(os.path.join(d["h2m"], f"{i}_hoi_mesh.npy"), T)
        hun = (s.hand_rest.numpy().astype(np.float64) - T[:3, 3]) @ np.linalg.inv(T[:3, :3]).T      # MoGe -> Hunyuan
        write_ply(os.path.join(d["mano"], f"{i}_hamer_aligned_mano.ply"), hun, s.hand_faces.numpy())
        write_ply(os.path.join(d["hun"], f"{i}_hoi_mesh.ply"), hun, s.hand_faces.numpy())
        np.save(os.path.join(d["hamer"], f"{i}_kps_for_guidance.npy"),
                {"mano_2d_kps": s.kps_2d.numpy() / 8.0, "mano_3d_kps": np.zeros((21, 3)), "cam_t": np.zeros(3)}, allow_pickle=True)
    jpath = os.path.join(root, "J_regressor_hamer.pt")
    torch.save(torch.from_numpy(standin_j_regressor(0)), jpath)
    return d, jpath


def _kwargs(d):
    return dict(project_root="/nonexistent", cropped_obj_img_dir=d["img"], mask_dir=d["mask"], moge_out_dir=d["moge"],
                hunyuan_hoi_mesh_dir=d["hun"], hamer_out_dir=d["hamer"], h2m_rt_dir=d["h2m"], aligned_mano_dir=d["mano"],
                guidance_out_dir=d["out"])


def test_signature_flags_and_file_names_match_the_reference():
    params = list(inspect.signature(R.run).parameters)
    assert params[:10] == REF_ARGS
    assert inspect.signature(R.run).parameters["task_list_file"].default is None
    src = inspect.getsource(R.main)
    for a in REF_ARGS:
        assert f'"--{a}"' in src
    p = R.index_paths("017_cropped_obj_1.png", "/a", "/m", "/g", "/h", "/k", "/t", "/n", "/o")
    assert p["index"] == "017"
    assert p["cropped_hand_mask_path"] == "/m/017_cropped_hand_mask.png" and p["cropped_obj_mask_path"] == "/m/017_cropped_obj_mask.png"
    assert p["moge_fov_path"] == "/g/017_cropped_hoi/fov.json" and p["T_h2m_path"] == "/t/017_hoi_mesh.npy"
    assert p["aligned_mano_mesh_path"] == "/n/017_hamer_aligned_mano.ply" and p["hunyuan_hoi_mesh_path"] == "/h/017_hoi_mesh.ply"
    assert p["hamer_for_guid_path"] == "/k/017_kps_for_guidance.npy"
    assert p["save_path_obj"] == "/o/017_obj.ply" and p["save_path_hand"] == "/o/017_hand.ply"


def test_no_model_fails_loudly_and_task_list_semantics(tmp_path, monkeypatch):
    from followmyhold_b200 import _lib
    d, _ = write_dataset(str(tmp_path), 1)
    monkeypatch.delenv(R.MODEL_ENV, raising=False)
    with pytest.raises(_lib.FohoLibraryError):
        R.run(**_kwargs(d))
    chunks = [["a_1.png", "b_1.png"], ["c_1.png"]]
    tl = tmp_path / "tasks.json"
    tl.write_text(json.dumps(chunks))
    monkeypatch.setenv("SLURM_ARRAY_TASK_ID", "1")
    assert R._load_task_list(str(tl), d["img"]) == ["c_1.png"]
    monkeypatch.delenv("SLURM_ARRAY_TASK_ID")
    monkeypatch.setenv("RANK", "0"); monkeypatch.setenv("WORLD_SIZE", "1")
    assert R._load_task_list(str(tl), d["img"]) == ["a_1.png", "b_1.png"]
    assert R._load_task_list(None, d["img"]) == ["000_cropped_obj_1.png"]


def test_inputs_loader_and_mock_surface(tmp_path):
    d, _ = write_dataset(str(tmp_path), 2, empty_mask_for=1)
    kw = _kwargs(d); kw.pop("project_root")
    p0 = R.index_paths("000_cropped_obj_1.png", **kw)
    inp = R.load_image_inputs(p0, 512, np.random.default_rng(0))
    assert inp["hw"] == (64, 64) and inp["fovx"] == 41.0 and inp["cloud"].shape == (512, 3) and inp["kps"].shape == (21, 2)
    s = make_guidance_sample(32, 2000, 200)
    assert np.allclose(inp["hand_moge"], s.hand_rest.numpy(), atol=1e-5)           # Hunyuan ply -> T_h2m -> MoGe
    p1 = R.index_paths("001_cropped_obj_1.png", **kw)
    assert R.load_image_inputs(p1, 512, np.random.default_rng(0)) == {"skip": "empty mask"}
    # fewer points than requested: repeated, not invented
    few = R.load_image_inputs(p0, 4096, np.random.default_rng(0))["cloud"]
    assert few.shape == (4096, 3) and len(np.unique(few, axis=0)) == 2000
    # the mock surface is closed: every edge of the blocky mesh is shared by exactly two faces
    m = R.MockGuidanceModel(D=16, latent_elems=1024)
    x = np.linspace(-1.1, 1.1, 16)
    X, Y, Z = np.meshgrid(x, x, x, indexing="ij")
    v, f = m.extract_mesh(np.sqrt(X * X + Y * Y + Z * Z) - 0.6)
    e = np.sort(np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]), axis=1)
    _, cnt = np.unique(e, axis=0, return_counts=True)
    assert len(v) > 0 and (cnt == 2).all()
    assert np.abs(np.linalg.norm(v, axis=1) - 0.6).max() < 0.3
    assert np.allclose(R.similarity_about(v, np.array([1, 0, 0, 0, 1, 0, 0, 0.0]), np.zeros(3)), v)


def test_nan_images_are_reported_like_the_reference(capsys):
    """NaN total: "Total loss is NaN" (pipelines.py:1443,1591); an image whose hand / object-only phase hit it is
    the reference's `return None` -> the stage's per-image error line, nothing written (run.py:141,257-259)."""
    paths = ["/d/000_cropped_obj_1.png", "/d/001_cropped_obj_1.png", "/d/002_cropped_obj_1.png"]
    skip = R.report_nan_images(paths, {10: {1: 0}, 14: {2: 7}}, failed={1})
    out = capsys.readouterr().out.splitlines()
    assert skip == {1}
    assert out == ["Total loss is NaN", "Error in processing 001_cropped_obj_1.png : cannot unpack non-iterable NoneType object",
                   "Total loss is NaN"]
    assert R.report_nan_images(paths, {}, failed=set()) == set() and capsys.readouterr().out == ""


@pytest.mark.gpu
def test_stage_writes_the_debug_dumps_when_asked(tmp_path, monkeypatch):
    """FOHO_DEBUG_DIR (pipelines.py:1076-1091): params.json + losses.txt per image, a line every 10th inner iteration."""
    from followmyhold_b200.guidance.config import OptimizationConfig
    d, jpath = write_dataset(str(tmp_path), 2)
    dbg = tmp_path / "debug"
    monkeypatch.setenv("FOHO_DEBUG_DIR", str(dbg))
    cfg = OptimizationConfig().with_steps(6)
    cfg.optimization_steps_hand, cfg.optimization_steps_scale, cfg.optimization_steps_joint = 12, 3, 11
    R.run(**_kwargs(d), model=R.MockGuidanceModel(D=32, latent_elems=1024), batch_size=2, n_cloud=1024, config=cfg,
          j_regressor_path=jpath)
    dirs = sorted(os.listdir(dbg))
    assert len(dirs) == 2 and dirs[0].endswith("_exp_obj000_inpainted") and dirs[1].endswith("_exp_obj001_inpainted")
    params = json.load(open(dbg / dirs[0] / "params.json"))
    assert params["optimization_steps_joint"] == 11 and params["phase1_hand_lrs"] == {"scale": 1e-2, "trans": 1e-2, "rot": 0.5}
    assert params["noise_obj_lr2"] == 1e-2 and params["use_intersection_loss"] is True
    lines = open(dbg / dirs[1] / "losses.txt").read().splitlines()
    # steps 2 (hand: k = 0, 10), 3 (object: k = 0), 4 and 5 (joint: k = 0, 10)
    assert [l.split(",")[0] + "," + l.split(",")[1] for l in lines] == [
        "Denoise step 2 phase 1, Opt step 0", "Denoise step 2 phase 1, Opt step 10", "Denoise step 3 phase 1.5, Opt step 0",
        "Denoise step 4 phase 2, Opt step 0", "Denoise step 4 phase 2, Opt step 10",
        "Denoise step 5 phase 2, Opt step 0", "Denoise step 5 phase 2, Opt step 10"]
    assert all("nan" not in l for l in lines)
    assert os.path.exists(os.path.join(d["out"], "000_obj.ply")) and os.path.exists(os.path.join(d["out"], "001_hand.ply"))


@pytest.mark.gpu
def test_stage_runs_three_synthetic_images_and_skips_done_ones(tmp_path, capsys):
    from followmyhold_b200.guidance.config import OptimizationConfig
    d, jpath = write_dataset(str(tmp_path), 3)
    cfg = OptimizationConfig().with_steps(6)
    cfg.optimization_steps_hand, cfg.optimization_steps_scale, cfg.optimization_steps_joint = 4, 3, 2
    model = R.MockGuidanceModel(D=32, latent_elems=1024)
    R.run(**_kwargs(d), model=model, batch_size=2, n_cloud=1024, config=cfg, j_regressor_path=jpath)
    for k in range(3):
        hand = load(os.path.join(d["out"], f"{k:03d}_hand.ply"))
        obj = load(os.path.join(d["out"], f"{k:03d}_obj.ply"))
        assert hand.vertices.shape == (778, 3) and hand.faces.shape == (1538, 3)
        assert obj.vertices.shape[0] > 100 and np.isfinite(obj.vertices).all()
        src = make_guidance_sample(32, 2000, 200 + k).hand_rest.numpy()
        moved = np.abs(hand.vertices - src).max()
        assert 0 < moved < 0.2                       # the hand phase moved the hand, but not far
    stamp = os.path.getmtime(os.path.join(d["out"], "000_hand.ply"))
    capsys.readouterr()
    R.run(**_kwargs(d), model=model, batch_size=2, n_cloud=1024, config=cfg, j_regressor_path=jpath)
    out = capsys.readouterr().out
    assert out.count("already exists, skipping") == 3 and os.path.getmtime(os.path.join(d["out"], "000_hand.ply")) == stamp


@pytest.mark.gpu
def test_stage_with_a_torch_decoder_model_equals_the_stand_in_path(tmp_path):
    """A model exposing ``decode`` (a differentiable torch decoder, as the reference's VAE would be wrapped) is driven
    through ``run_schedule_decoder``; with the linear stand-in expressed in torch the stage writes the same meshes."""
    from followmyhold_b200.guidance.config import OptimizationConfig

    class TorchDecoderModel(R.MockGuidanceModel):
        def decode(self, x1):
            flat0 = self._sdf0.reshape(x1.shape[0], -1)
            tap = self.tap.to(x1.device)
            return flat0.scatter(1, tap.view(1, -1).expand(x1.shape[0], -1), flat0[:, tap] + self.alpha * x1).reshape(self._sdf0.shape)

    cfg = OptimizationConfig().with_steps(6)
    cfg.optimization_steps_hand, cfg.optimization_steps_scale, cfg.optimization_steps_joint = 4, 3, 2
    outs = []
    for k, cls in enumerate((R.MockGuidanceModel, TorchDecoderModel)):
        d, jpath = write_dataset(str(tmp_path / f"run{k}"), 2)
        R.run(**_kwargs(d), model=cls(D=32, latent_elems=1024), batch_size=2, n_cloud=1024, config=cfg, j_regressor_path=jpath)
        outs.append([(load(os.path.join(d["out"], f"{i:03d}_hand.ply")).vertices, load(os.path.join(d["out"], f"{i:03d}_obj.ply")).vertices)
                     for i in range(2)])
    for (h0, o0), (h1, o1) in zip(*outs):
        assert np.allclose(h0, h1, atol=2e-5)
        assert o0.shape == o1.shape and np.allclose(o0, o1, atol=2e-4)
