"""Shim behind ``foho.guidance.run`` when ``followmyhold_b200.dropin`` is installed: the mirror's public names, and
its CLI when run as ``python -m foho.guidance.run`` (what ``foho.main`` does, src/foho/main.py:80-91)."""
from followmyhold_b200.guidance.run import *  # noqa: F401,F403
from followmyhold_b200.guidance.run import main, run  # noqa: F401

if __name__ == "__main__":
    main()
