"""Export the reference's parameter catalogue (data, not code) to a compact JSON table.

Run in the build container: reads /root/reference/src/ska_sdp_exec_swiftly/swift_configs.py
(a dict of SwiftlyConfig keyword sets) and writes swift_configs.json next to the package's
swift_configs.py as rows [name, W, fov, N, Nx, yB, yN, yP, xA, xM].
"""
import json
import os
import runpy

SRC = "/root/reference/src/ska_sdp_exec_swiftly/swift_configs.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                   "ska_sdp_distributed_fourier_transform_b200", "swift_configs.json")
cfgs = runpy.run_path(SRC)["SWIFT_CONFIGS"]
cols = ["W", "fov", "N", "Nx", "yB_size", "yN_size", "yP_size", "xA_size", "xM_size"]
rows = [[name] + [c.get(k) for k in cols] for name, c in cfgs.items()]
with open(OUT, "w") as f:
    json.dump({"columns": ["name"] + cols, "rows": rows}, f, separators=(",", ":"))
print(len(rows), "entries ->", OUT)
