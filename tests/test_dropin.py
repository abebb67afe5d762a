"""The stage mirrors drop into ``foho.main`` unchanged: ``foho.main`` runs ``python3 -m foho.<stage>`` with the
reference's ``src`` FIRST on PYTHONPATH (src/foho/main.py:24-27,80-91), so the redirect is an import hook
(followmyhold_b200.dropin), exercised here against a stand-in ``foho`` package laid out like the reference's."""
import os
import subprocess
import sys

import pytest

from followmyhold_b200 import dropin

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def fake_reference(tmp_path):
    src = tmp_path / "src"
    for pkg in ("foho", "foho/alignment", "foho/guidance", "foho/utils"):
        (src / pkg).mkdir(parents=True)
        (src / pkg / "__init__.py").write_text("")
    for mod in ("alignment/h2m", "alignment/mano", "guidance/run", "utils/runner"):
        (src / "foho" / (mod + ".py")).write_text(
            "def run(*a, **k):\n    return 'REFERENCE'\n\ndef main():\n    print('REFERENCE " + mod + "')\n\n"
            "if __name__ == '__main__':\n    main()\n")
    site = tmp_path / "site"
    site.mkdir()
    (site / "sitecustomize.py").write_text(dropin.PTH_LINE + "\n")       # what the .pth line does at start-up
    return src, site


def _run(args, src, site, **env):
    e = dict(os.environ)
    e["PYTHONPATH"] = f"{src}:{site}:{REPO}"                              # the reference's src first, like foho.main
    e.update(env)
    return subprocess.run([sys.executable] + args, capture_output=True, text=True, env=e, timeout=300)


def test_python_dash_m_runs_the_mirror_and_leaves_other_modules_alone(fake_reference):
    src, site = fake_reference
    r = _run(["-m", "foho.alignment.h2m", "--help"], src, site)
    assert r.returncode == 0 and "REFERENCE" not in r.stdout
    for flag in ("--hunyuan_mesh_dir", "--moge_out_dir", "--h2m_rt_dir"):                  # h2m.py:63-70
        assert flag in r.stdout
    r = _run(["-m", "foho.utils.runner"], src, site)                                       # not a redirected module
    assert r.stdout.strip() == "REFERENCE utils/runner"
    r = _run(["-m", "foho.alignment.h2m"], src, site, FOHO_B200_DROPIN="0")                # switch: reference stage
    assert r.stdout.strip() == "REFERENCE alignment/h2m"
    code = ("import foho.guidance.run as g, foho.alignment.mano as m, foho.utils.runner as u;"
            "print(g.run.__module__, m.run.__module__, u.run())")
    r = _run(["-c", code], src, site, FOHO_B200_DROPIN_GUIDANCE="1")
    assert r.stdout.split() == ["followmyhold_b200.guidance.run", "followmyhold_b200.alignment.mano", "REFERENCE"], r.stderr
    # the guidance stage optimises a different energy from the reference's: it is redirected on opt-in only
    r = _run(["-c", code], src, site)
    assert r.stdout.split() == ["foho.guidance.run", "followmyhold_b200.alignment.mano", "REFERENCE"], r.stderr


def test_install_is_idempotent_and_pth_file(tmp_path):
    before = list(sys.meta_path)
    try:
        assert dropin.install() and dropin.install()
        assert sum(isinstance(f, dropin.StageRedirect) for f in sys.meta_path) == 1
        f = dropin.StageRedirect()
        assert f.find_spec("foho.main") is None and f.find_spec("foho.guidance") is None
        assert f.find_spec("foho.guidance.run") is None                   # opt-in only
        os.environ["FOHO_B200_DROPIN_GUIDANCE"] = "1"
        spec = f.find_spec("foho.guidance.run")
        assert spec is not None and spec.origin.endswith("_shims/foho_guidance_run.py") and os.path.exists(spec.origin)
        os.environ["FOHO_B200_DROPIN"] = "0"
        assert f.find_spec("foho.guidance.run") is None and dropin.install() is False
    finally:
        os.environ.pop("FOHO_B200_DROPIN", None)
        os.environ.pop("FOHO_B200_DROPIN_GUIDANCE", None)
        dropin.uninstall()
        assert sys.meta_path == before
    p = dropin.write_pth(str(tmp_path))
    assert os.path.basename(p) == "foho_b200_dropin.pth" and open(p).read().strip() == dropin.PTH_LINE
    for name, target in dropin.REDIRECTS.items():
        assert os.path.exists(os.path.join(REPO, "followmyhold_b200", "_shims", name.replace(".", "_") + ".py"))
        assert os.path.exists(os.path.join(REPO, *target.split(".")) + ".py")
