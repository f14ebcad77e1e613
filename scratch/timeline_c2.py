import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from pero_pretraining_b200 import _lib, ops
L = _lib.lib(); dev = torch.device("cuda:0")
stream = torch.cuda.current_stream().cuda_stream
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
names = ["mma_top", "tempty_ok", "first_full", "issued", "epi_tfull", "epi_done", "prod_first", "prod_last"]
N, K, D = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (8192, 8192, 256))]
w = torch.randn(K, D, device=dev); xb = torch.randn(N, D, device=dev).bfloat16()
cb = ops.PreparedCodebook(K, D, dev).prepare(w)
packed = torch.empty(N, dtype=torch.int64, device=dev)
tl = torch.zeros(8192 + 148 * 4 + 64, dtype=torch.int64, device=dev)
for _ in range(3):
    tl.zero_(); L.pero_vq_packed_init(packed.data_ptr(), N, stream); flush.zero_()
    L.pero_debug_set_timeline(tl.data_ptr(), tl.numel() // 8192)       # every launch advances the debug pointer by one 64 KiB slot
    _lib.check(L.pero_vq_assign_bf16(xb.data_ptr(), N, K, D, cb.blob.data_ptr(), 0, packed.data_ptr(), stream), "assign")
    torch.cuda.synchronize()
    L.pero_debug_set_timeline(None, 0)
t = tl[:4096].view(512, 8).cpu(); t0 = int(t[0][6])
c = tl[4096:4096 + 4].cpu()
print("cta0 ns: setup", int(c[1]-c[0]), "epi_done", int(c[2]-c[0]), "exit", int(c[3]-c[0]))
print("unit " + " ".join(f"{n:>11s}" for n in names))
for u in range(15):
    if int(t[u].max()) == 0: break
    print(f"{u:4d} " + " ".join(f"{(int(v)-t0) if int(v) else -1:11d}" for v in t[u]))
