import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import leann_rs_b200 as P
n, d = int(sys.argv[1]), 768
torch.manual_seed(0)
W = torch.randn(32, d, device='cuda')
def gen(m):
    x = torch.randn(m, 32, device='cuda') @ W + 0.3 * torch.randn(m, d, device='cuda')
    return torch.nn.functional.normalize(x, dim=1).contiguous()
x = gen(n); q = gen(10000)
torch.cuda.synchronize(); t = time.time()
s = P.HnswSearcher.build(x, 32, 64)
torch.cuda.synchronize(); print('build', n, 'in %.2fs' % (time.time() - t), s.info())
flat = P.FlatSearcher.from_vectors(x, metric=P.METRIC_IP)
t = time.time(); gk, gd, gc = flat.search_device(q, 10, 0); torch.cuda.synchronize(); print('exact %.3fs' % (time.time() - t))
for ef in (64, 128, 256):
    stats = torch.zeros((q.shape[0], 4), dtype=torch.int64, device='cuda')
    s.search_device(q, 10, ef, stats=stats); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); k, dd, c = s.search_device(q, 10, ef); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    rec = (k.unsqueeze(2) == gk.unsqueeze(1)).any(2).float().mean().item()
    st = stats.float().mean(0).tolist()
    byts = stats[:, 0].sum().item() * d * 4 + stats[:, 1].sum().item() * 64 * 4 + stats[:, 2].sum().item() * 32 * 4
    print('ef', ef, 'recall %.4f' % rec, '%.2f ms' % ms, 'QPS %.0f' % (q.shape[0] / ms * 1e3), 'ndist %.0f hops %.0f' % (st[0], st[1]), 'GB/s %.0f' % (byts / ms / 1e6))
