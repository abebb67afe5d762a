"""Drop the three stage mirrors into ``foho.main`` without editing the reference.

``foho.main`` launches every stage as ``python3 -m foho.<module> --flag value`` in a subprocess whose
``PYTHONPATH`` starts with the reference's own ``src`` (``src/foho/main.py:21,24-27,80-91``;
``src/foho/utils/runner.py:10-21``), so a shadowing package later on the path never wins.  What does work
without touching the reference is an import hook: ``install()`` puts a finder in front of ``sys.meta_path``
that resolves exactly three module names

    foho.alignment.h2m   foho.alignment.mano   foho.guidance.run

to thin shims (``_shims/``) that re-export the mirrors' ``run`` / ``main`` and call ``main()`` when executed
by ``python -m``; every other ``foho.*`` module (``foho.main``, the configs, the upstream stages) still comes
from the reference.  To have it in every subprocess put one line into a ``.pth`` file of the environment the
stages run in (``write_pth()`` does that) -- or import this module from a ``sitecustomize.py``:

    import followmyhold_b200.dropin as d; d.install()

Side effect to know about: importing the package sets ``CUDA_DEVICE_MAX_CONNECTIONS=32`` if it is unset
(``followmyhold_b200/__init__.py``), so with the ``.pth`` line every interpreter of that environment starts
with 32 hardware work queues instead of 8; export the variable yourself to keep another value.

``FOHO_B200_DROPIN=0`` switches the redirect off (A/B runs against the reference stages).

The two alignment stages are redirected by default: they compute what the reference computes (golden-pinned
ICP).  The guidance stage is redirected ONLY with ``FOHO_B200_DROPIN_GUIDANCE=1``: its energy is not the
reference's yet -- the rendered normal / disparity / silhouette terms (pipelines.py:1339-1349,1566-1588) are
absent, chamfer and volume terms stand in for them -- so its outputs are not comparable with the reference's
although the file names match, and swapping it silently would be wrong.  It additionally needs the Hunyuan3D
networks wrapped in a ``GuidanceModel`` named by ``FOHO_B200_GUIDANCE_MODEL`` (guidance/run.py); without one
it fails loudly.
"""
from __future__ import annotations

import importlib.abc
import importlib.util
import os
import sys
from typing import Optional

REDIRECTS = {
    "foho.alignment.h2m": "followmyhold_b200.alignment.h2m",        # src/foho/alignment/h2m.py:12-72
    "foho.alignment.mano": "followmyhold_b200.alignment.mano",      # src/foho/alignment/mano.py:12-61
    "foho.guidance.run": "followmyhold_b200.guidance.run",          # src/foho/guidance/run.py:178-289
}
_SHIM_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_shims")
PTH_NAME = "foho_b200_dropin.pth"
PTH_LINE = "import followmyhold_b200.dropin as _foho_b200_dropin; _foho_b200_dropin.install()"


OPT_IN = {"foho.guidance.run": "FOHO_B200_DROPIN_GUIDANCE"}       # redirected only when this variable is truthy
_OFF = ("0", "false", "False", "")


def enabled(fullname: Optional[str] = None) -> bool:
    if os.environ.get("FOHO_B200_DROPIN", "1") in _OFF:
        return False
    var = OPT_IN.get(fullname) if fullname else None
    return var is None or os.environ.get(var, "0") not in _OFF


class StageRedirect(importlib.abc.MetaPathFinder):
    """Resolves the three stage modules to their shims; declines everything else."""

    def find_spec(self, fullname, path=None, target=None):
        if fullname not in REDIRECTS or not enabled(fullname):
            return None
        shim = os.path.join(_SHIM_DIR, fullname.replace(".", "_") + ".py")
        return importlib.util.spec_from_file_location(fullname, shim)


def install() -> bool:
    """Idempotent.  Returns True when the redirect is active in this interpreter."""
    if not enabled():
        return False
    if not any(isinstance(f, StageRedirect) for f in sys.meta_path):
        sys.meta_path.insert(0, StageRedirect())
    return True


def uninstall() -> None:
    sys.meta_path[:] = [f for f in sys.meta_path if not isinstance(f, StageRedirect)]


def write_pth(site_dir: Optional[str] = None) -> str:
    """Write ``foho_b200_dropin.pth`` into ``site_dir`` (default: this interpreter's site-packages) so every
    interpreter of that environment installs the redirect at start-up.  ``followmyhold_b200`` itself must be
    importable there (repo root on ``PYTHONPATH`` or in a ``.pth`` of its own).  Returns the file written."""
    if site_dir is None:
        import site
        site_dir = site.getsitepackages()[0]
    path = os.path.join(site_dir, PTH_NAME)
    with open(path, "w") as f:
        f.write(PTH_LINE + "\n")
    return path
