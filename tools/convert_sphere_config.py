#!/usr/bin/env python
"""Convert the reference's per-link collision-sphere YAML files (data/configs/*/..._sphere_config.yaml: link name ->
list of [x, y, z, radius] in the link frame) into the flat JSON the package ships:
    {"spheres": [{"link": name, "offset": [x, y, z], "radius": r}, ...]}
usage: python tools/convert_sphere_config.py [src_configs_dir] [dst_configs_dir]"""
import json
import sys
from pathlib import Path

import yaml


def main():
    here = Path(__file__).resolve().parent.parent
    src = Path(sys.argv[1]) if len(sys.argv) > 1 else Path("/root/reference/torch_robotics/data/configs")
    dst = Path(sys.argv[2]) if len(sys.argv) > 2 else here / "torch_robotics_b200" / "data" / "configs"
    for f in sorted(src.rglob("*_sphere_config.yaml")):
        cfg = yaml.safe_load(f.read_text())
        scalars = {k: v for k, v in cfg.items() if not isinstance(v, list)}        # e.g. tiago's `margin`
        spheres = []
        for link, lst in cfg.items():
            for s in (lst if isinstance(lst, list) else []):
                rec = {"link": link, "offset": [float(v) for v in s[:3]], "radius": float(s[3])}
                if len(s) > 4:
                    rec["group"] = int(s[4])
                spheres.append(rec)
        out = dst / f.relative_to(src).with_suffix(".json")
        out.parent.mkdir(parents=True, exist_ok=True)
        out.write_text(json.dumps({"source": f.name, **scalars, "spheres": spheres}, indent=1) + "\n")
        print("wrote", out, len(spheres), "spheres")


if __name__ == "__main__":
    main()
