"""Per-kernel SASS evidence (profiles/sass/): which Blackwell-specific instructions each hot kernel contains, with short
excerpts around them.  Runs on the build box (no GPU): `python scripts/sass_evidence.py` after `build()`."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "followmyhold_b200", "libfoho_b200.so")
OUT = os.path.join(ROOT, "profiles", "sass")
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCATOMSWS", "USETMAXREG", "SYNCS", "MUFU.EX2", "HMMA", "F2FP",
        "FENCE.VIEW.ASYNC", "ELECT"]
WANT = {"k_gemm_tc": "decoder_gemm", "k_attn_fwd2": "decoder_attn_v2", "k_attn_fwd": "decoder_attn_v1", "k_stream_tma": "guidance_stream_tma",
        "k_chamfer_c2h_walk": "chamfer_walk", "k_voxdist_staged": "voxdist_staged", "k_icp_step": "icp_step", "k_icp_loop": "icp_loop",
        "k_attn_bwd": "decoder_attn_bwd", "k_rs_raster": "raster", "k_dmc_verts": "dmc_verts", "k_rc_mask": "remove_close"}


def main():
    os.makedirs(OUT, exist_ok=True)
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    summary, seen = [], set()
    for f in funcs[1:]:
        name = f.split("\n", 1)[0].strip()
        short = None
        for k in sorted(WANT, key=len):                       # longest match wins (k_attn_fwd2 over k_attn_fwd)
            if re.search(r"\d+" + k + r"(?![a-z0-9_])", name):
                short = k
        if short is None:
            continue
        lines = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4,6}\*/", l)]
        ops = collections.Counter()
        for l in lines:
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m and any(m.group(1).startswith(k) for k in KEYS):
                ops[m.group(1)] += 1
        summary.append((short, name, len(lines), dict(ops)))
        tag = WANT[short]
        if short == "k_gemm_tc" and "Li256ELb0ELb0" not in name:     # one excerpt: 256-wide tile, both operands K-major
            continue
        if tag in seen:
            continue
        seen.add(tag)
        idxs = [i for i, l in enumerate(lines) if re.search(r"UTCHMMA|UTMALDG|UBLKCP|LDTM|STTM|UTCBAR|USETMAXREG", l)][:8]
        with open(os.path.join(OUT, f"r02_{tag}.sass.txt"), "w") as fh:
            fh.write(f"# cuobjdump -sass followmyhold_b200/libfoho_b200.so (sm_100a), function {name}\n"
                     f"# {len(lines)} instructions; Blackwell-specific mnemonics: {dict(ops)}\n")
            last = -100
            for i in idxs:
                if i - last < 8:
                    continue
                fh.write(f"\n# --- around instruction {i}\n")
                for l in lines[max(0, i - 6):i + 8]:
                    fh.write(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip()) + "\n")
                last = i
    with open(os.path.join(OUT, "r02_INDEX.md"), "w") as fh:
        fh.write("# SASS evidence (round 2): `cuobjdump -sass followmyhold_b200/libfoho_b200.so`, sm_100a\n\n"
                 "Per kernel: instruction count and the Blackwell-specific mnemonics it contains (B200_PROFILING.md: `tcgen05.mma` = "
                 "`UTCHMMA`, `tcgen05.commit` = `UTCBAR`, `tcgen05.ld/st` = `LDTM/STTM`, TMA = `UTMALDG` (tensor) / `UBLKCP` (1-D bulk), "
                 "`tcgen05.alloc` = `UTCATOMSWS`, `setmaxnreg` = `USETMAXREG`).  Excerpts around those instructions are in the "
                 "`r02_*.sass.txt` files beside this one; regenerate with `python scripts/sass_evidence.py`.\n\n"
                 "| kernel | mangled name | instructions | mnemonics |\n|---|---|---|---|\n")
        for short, name, n, ops in summary:
            fh.write(f"| `{short}` | `{name[:100]}` | {n} | {ops} |\n")
    print(f"wrote {len(seen)} excerpts and the index to {OUT}")


if __name__ == "__main__":
    main()
