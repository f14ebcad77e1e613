"""e2e leg of bench.py under the PERO_E2E_* switches (host-side cost study)."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import bench
dev = torch.device("cuda:0"); torch.cuda.set_device(0)
bench.bind_to_gpu_numa_node(0)
batch = bench.make_batch(0)
for graph, bf16 in ((1, 1), (0, 1), (1, 0), (0, 0), (1, 1)):
    os.environ["PERO_E2E_GRAPH"] = str(graph); os.environ["PERO_E2E_BF16"] = str(bf16)
    med, mn, allt, h2d, d2h, loss = bench.e2e_leg(batch, dev, False, 30, 5)
    print(f"graph={graph} bf16={bf16}: median {med*1e3:.3f} ms min {mn*1e3:.3f} all {[round(t*1e3,3) for t in allt]} h2d {h2d/1e6:.1f} MB loss {loss:.4f}", flush=True)
