"""How long does opening an index take (HnswSearcher::load, hnsw.rs:18-75: read the whole `.index`, build the device layout)?
  python benchmarks/load_probe.py [n] [d]"""
import json, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, leann_rs_b200 as P
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 768
dev = torch.device("cuda", 0)
g = torch.Generator(device=dev); g.manual_seed(1)
W = torch.randn((32, d), generator=g, device=dev)
x = torch.cat([torch.nn.functional.normalize(torch.randn((1 << 18, 32), generator=g, device=dev) @ W + 0.3 * torch.randn((1 << 18, d), generator=g, device=dev), dim=1)
               for _ in range((n + (1 << 18) - 1) >> 18)])[:n].contiguous()
t0 = time.time(); s = P.HnswSearcher.build(x, 32, 64); torch.cuda.synchronize(); t_build = time.time() - t0
td = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
base = os.path.join(td, "documents.leann")
t0 = time.time(); s.save(base); t_save = time.time() - t0
size = os.path.getsize(base.replace(".leann", ".index"))
s.close()
ts, tc = [], []
for _ in range(3):
    t0 = time.time(); s2 = P.HnswSearcher.load(base, d); ts.append(time.time() - t0); s2.close()
s2 = P.HnswSearcher.load(base, d); t0 = time.time(); s2.write_layout_cache(base); t_cache = time.time() - t0; s2.close()
for _ in range(3):
    t0 = time.time(); s2 = P.HnswSearcher.load(base, d); tc.append(time.time() - t0); assert s2.layout_cache_used; s2.close()
print(json.dumps({"bench": "load", "n": n, "d": d, "file_GB": round(size / 1e9, 2), "build_s": round(t_build, 2), "save_s": round(t_save, 2),
                  "open_s_parse": [round(t, 3) for t in ts], "open_s_layout_cache": [round(t, 3) for t in tc], "write_cache_s": round(t_cache, 2),
                  "cache_GB": round(os.path.getsize(base.replace(".leann", ".cuda-layout")) / 1e9, 2),
                  "open_GBps": round(size / 1e9 / min(ts + tc), 2), "io_threads": os.environ.get("LEANN_CUDA_IO_THREADS", "default min(8, cores)")}))
import shutil; shutil.rmtree(td)
