#!/usr/bin/env python
"""Generate tests/golden/ref_golden_config.json by EXECUTING the reference's config loader
(src/foho/configs/pipeline.py ``load_config``) on the .env texts below.  Authoring container only:

    PYTHONPATH=/root/reference/src python tests/golden/make_golden_config.py
"""
import dataclasses
import json
import os
import tempfile

from foho.configs.pipeline import load_config

CASES = {
    "defaults": '# comment\nPROJECT_ROOT="/proj/FOHO"\nBASE_DIR="/data/out"\nIMAGE_PATH="/x/example.png"\nCONDA_SH="/c/conda.sh"\n',
    "overrides": ("PROJECT_ROOT = '/p'\n\nBASE_DIR=/b\nSPLIT_PATH=/s.csv\nCONDA_SH=/c.sh\nH2M_RT_PATH = \"/custom/h2m\"\n"
                  "GUIDANCE_OUT_PATH='/g out'\nMOGE_OUT_PATH=\"'/quoted'\"\nNOT A KEY VALUE LINE\n#HAMER_OUT_PATH=/ignored\n"
                  "ALIGNED_MANO_PATH=/a=b\nMASK_DIR_PATH=\n"),
}
FIELDS = ["project_root", "base_dir", "cropped_inpainted_obj", "mask_dir_path", "moge_out_path", "hunyuan_hoi_mesh_path",
          "hamer_out_path", "h2m_rt_path", "aligned_mano_path", "guidance_out_path"]

out = {}
for name, text in CASES.items():
    with tempfile.NamedTemporaryFile("w", suffix=".env", delete=False) as f:
        f.write(text)
    cfg = dataclasses.asdict(load_config(f.name))
    os.unlink(f.name)
    out[name] = {"env_text": text, "config": {k: cfg[k] for k in FIELDS}}
here = os.path.dirname(os.path.abspath(__file__))
json.dump(out, open(os.path.join(here, "ref_golden_config.json"), "w"), indent=1)
print(json.dumps(out["overrides"]["config"], indent=1))
