import sys, statistics, torch
sys.path.insert(0, "/root/repo")
import hls_designs_b200 as hd
dev = torch.device("cuda", 0)
for n, batch in ((256, 4096), (512, 1024), (1024, 256)):
    g = torch.Generator(device=dev).manual_seed(1)
    A = torch.rand((batch, n, n), generator=g, device=dev); A = 0.5 * (A + A.transpose(1, 2)); A.diagonal(dim1=1, dim2=2).add_(float(n)); A = A.contiguous()
    for label, kw, b in (("blocked", dict(design="cholesky_v2.2"), batch), ("global", dict(design="cholesky_v2.2", path="global"), min(batch, 64))):
        X = A[:b].contiguous()
        hd.cholesky(X, **kw); torch.cuda.synchronize()
        ms = []
        for _ in range(3):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record(); hd.cholesky(X, **kw); e.record(); torch.cuda.synchronize(); ms.append(s.elapsed_time(e))
        t = statistics.median(ms)
        print(f"cholesky_v2.2 n={n} batch={b} {label:8s} {t:9.3f} ms {b / t * 1e3:10.4g} fact/s", flush=True)
