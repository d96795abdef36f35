import sys, json, types, torch
sys.path.insert(0, "/root/repo"); import bench
sys.argv=["x"]
args = bench.parse()
print(json.dumps(bench.stress_c5_block(args, torch.device("cuda", 0)))[:1500])
