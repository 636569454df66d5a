"""Where does the host-buffer call spend its time? Compares leann_cuda_search (pinned host buffers) with the
device-pointer call on the same batch."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, leann_rs_b200 as P
n, d, nq, k, ef = int(sys.argv[1]), 768, 10000, 10, int(sys.argv[2])
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1234)
W = torch.randn((32, d), generator=g, device=dev)
gen = lambda m: torch.nn.functional.normalize(torch.randn((m, 32), generator=g, device=dev) @ W + 0.3 * torch.randn((m, d), generator=g, device=dev), dim=1)
x = torch.cat([gen(1 << 18) for _ in range((n + (1 << 18) - 1) >> 18)])[:n].contiguous(); q = gen(nq)
idx = P.HnswSearcher.build(x, 32, 64)
hq = q.cpu().pin_memory(); hk = torch.empty((nq, k), dtype=torch.int64).pin_memory(); hd = torch.empty((nq, k)).pin_memory(); hc = torch.empty((nq,), dtype=torch.int32).pin_memory()
L = P.lib(); err = C.create_string_buffer(1024)
for i in range(3): idx.search_device(q, k, ef)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); [idx.search_device(q, k, ef) for _ in range(5)]; e1.record(); torch.cuda.synchronize()
print("device-resident ms/batch %.2f" % (e0.elapsed_time(e1) / 5))
for i in range(6):
    t0 = time.perf_counter()
    rc = L.leann_cuda_search(idx._h, C.c_void_p(hq.data_ptr()), nq, k, ef, None, 0, C.c_void_p(hk.data_ptr()), C.c_void_p(hd.data_ptr()), C.c_void_p(hc.data_ptr()), err, 1024)
    print("host call %d: %.2f ms rc=%d" % (i, (time.perf_counter() - t0) * 1e3, rc))
t0 = time.perf_counter(); dq = hq.to(dev, non_blocking=True); torch.cuda.synchronize(); print("H2D 30.7MB pinned: %.2f ms" % ((time.perf_counter() - t0) * 1e3))
