"""Kernel-only timing of the routing kernels on ONE GPU: ppcsr_bin_to_peers with all `parts` peer pointers aimed at
local buffers (same stores, no NVLink), CUDA events, for profiling under ncu.  python profiles/route_micro.py [parts]"""
import ctypes as C, importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pp = importlib.import_module("parallel-packed-csr_b200"); synth = importlib.import_module("parallel-packed-csr_b200.synth")
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 8
scale, B = 23, 10_000_000
dev = torch.device("cuda", 0); L = pp.load_library()
us, ud = synth.uniform(scale, 0, B, 7, device=dev); us, ud = us.to(torch.int32), ud.to(torch.int32)
starts = torch.tensor([p * (1 << scale) // parts for p in range(parts)] + [1 << scale], dtype=torch.int64, device=dev)
rec = [torch.empty(parts * B, dtype=torch.int64, device=dev) for _ in range(parts)]   # one receive buffer per "rank"
cnt = [torch.zeros(parts, dtype=torch.int64, device=dev) for _ in range(parts)]
u64 = C.c_uint64 * parts
rp, cp = u64(*[r.data_ptr() for r in rec]), u64(*[c.data_ptr() for c in cnt])
st = torch.cuda.current_stream().cuda_stream
def run():
    rc = L.ppcsr_bin_to_peers(0, st, starts.data_ptr(), parts, 0, us.data_ptr(), ud.data_ptr(), None, B, rp, None, cp, B)
    assert rc == 0, L.ppcsr_last_error()
for _ in range(3): run()
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"bin_to_peers parts={parts} B={B}: {e0.elapsed_time(e1)/10*1e3:.1f} us per batch; counts {[int(c[0]) for c in cnt]}")
