"""C3 with the reference's default masks (mean fill), in-launch fills on / off (B200FE_INLAUNCH_FILLS), for the library in B200FE_LIB."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import bench_configs as bc
r = bc.run_c3(torch.device("cuda:0"), steps=20, warmup=3, variants=("global", "mean"))
print(json.dumps({"inlaunch_fills": os.environ.get("B200FE_INLAUNCH_FILLS", "0"), "fill_lag": os.environ.get("B200FE_FILL_LAG", "600"),
                  "global_ms": r["variants"]["global"]["ms_per_step"], "mean_ms": r["variants"]["mean"]["ms_per_step"]}))
