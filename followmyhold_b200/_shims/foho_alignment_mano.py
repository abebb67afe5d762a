"""Shim behind ``foho.alignment.mano`` when ``followmyhold_b200.dropin`` is installed: the mirror's public names, and
its CLI when run as ``python -m foho.alignment.mano`` (what ``foho.main`` does, src/foho/main.py:80-91)."""
from followmyhold_b200.alignment.mano import *  # noqa: F401,F403
from followmyhold_b200.alignment.mano import main, run  # noqa: F401

if __name__ == "__main__":
    main()
