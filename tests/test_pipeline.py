"""In-process runner of the three replaced stages (followmyhold_b200.pipeline): config parsing and directory
derivation against golden output of the reference's own ``load_config`` (tests/golden/make_golden_config.py),
stage order and arguments against ``run_pipeline`` (src/foho/main.py:229-278)."""
import inspect
import json
import os

import pytest

from followmyhold_b200 import pipeline as PL

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "ref_golden_config.json")))


@pytest.mark.parametrize("case", sorted(G))
def test_directories_equal_the_reference_config_loader(case, tmp_path):
    p = tmp_path / "pipeline.env"
    p.write_text(G[case]["env_text"])
    sp = PL.load_stage_paths(str(p))
    for k, v in G[case]["config"].items():
        assert getattr(sp, k) == v, k


def test_errors_like_the_reference(tmp_path):
    with pytest.raises(FileNotFoundError):
        PL.load_stage_paths(str(tmp_path / "missing.env"))
    p = tmp_path / "bad.env"
    p.write_text("PROJECT_ROOT=/p\n")
    with pytest.raises(ValueError, match="PROJECT_ROOT and BASE_DIR are required"):
        PL.load_stage_paths(str(p))


def test_stage_order_arguments_and_working_directory(tmp_path, monkeypatch):
    from followmyhold_b200.alignment import h2m, mano
    from followmyhold_b200.guidance import run as grun
    proj = tmp_path / "proj"; proj.mkdir()
    env = tmp_path / "pipeline.env"
    env.write_text(f'PROJECT_ROOT="{proj}"\nBASE_DIR="{tmp_path}/out"\n')
    sp = PL.load_stage_paths(str(env))
    calls = PL.stage_calls(sp)
    assert [c[0] for c in calls] == ["alignment.h2m", "alignment.mano", "guidance.run"]          # main.py:229-278
    for (name, kw), mod in zip(calls, (h2m, mano, grun)):
        params = inspect.signature(mod.run).parameters
        assert set(kw) <= set(params), name                                                      # every kwarg exists
        required = [n for n, q in params.items() if q.default is inspect.Parameter.empty and q.kind != q.KEYWORD_ONLY]
        assert set(required) <= set(kw), name                                                    # and none is missing
    assert calls[2][1]["cropped_obj_img_dir"] == f"{tmp_path}/out/ours_inpaint" and calls[2][1]["mask_dir"] == f"{tmp_path}/out/cropped_hand_masks"
    seen = []
    monkeypatch.setattr(h2m, "run", lambda **k: seen.append(("h2m", os.getcwd(), k)))
    monkeypatch.setattr(mano, "run", lambda **k: seen.append(("mano", os.getcwd(), k)))
    monkeypatch.setattr(grun, "run", lambda **k: seen.append(("guidance", os.getcwd(), k)))
    here = os.getcwd()
    PL.run_hot_stages(str(env), model="M", device="cuda:3", batch_size=4)
    assert [s[0] for s in seen] == ["h2m", "mano", "guidance"]
    assert all(os.path.samefile(s[1], proj) for s in seen) and os.getcwd() == here               # cwd = project root, restored
    assert seen[0][2]["device"] == "cuda:3" and seen[2][2]["model"] == "M" and seen[2][2]["batch_size"] == 4
    for d in (sp.h2m_rt_path, sp.aligned_mano_path, sp.guidance_out_path):
        assert os.path.isdir(d)
