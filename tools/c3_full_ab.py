"""C3 full-`specaug` step (fbank + global CMVN -> time warp -> masks) for the library selected with B200FE_LIB."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools import bench_configs as bc
r = bc.run_c3(torch.device("cuda:0"), steps=20, warmup=3, variants=("global", "full"))
print(json.dumps({"lib": os.path.basename(os.environ.get("B200FE_LIB", "default")), "global_ms": r["variants"]["global"]["ms_per_step"], "full_ms": r["variants"]["full"]["ms_per_step"]}))
