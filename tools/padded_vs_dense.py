import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, wg_loader
wg = wg_loader.load()
def timeit(fn, iters=40):
    for i in range(4): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters
for cin, cout in ((512,128),(128,512),(1024,256),(256,1024)):
    rs = np.random.RandomState(0)
    layer = wg.Conv1x1Bn((rs.rand(cin, cout)-0.5).astype(np.float32), rs.rand(cout).astype(np.float32), rs.rand(cout).astype(np.float32), False)
    n, sets = 256, 3
    xs = [torch.rand((n,196,cin), device="cuda")-0.5 for _ in range(sets)]
    yd = [torch.empty((n,196,cout), device="cuda") for _ in range(sets)]
    yp = [torch.empty((n,16,16,cout), device="cuda") for _ in range(sets)]
    a = timeit(lambda i: layer(xs[i%sets], out=yd[i%sets]))
    b = timeit(lambda i: layer(xs[i%sets], out=yp[i%sets], out_padded=True))
    print(f"{cin}->{cout}: dense (TMA tensor stores) {a:.2f} us   padded frame (STG.128 rows, 30% more bytes) {b:.2f} us")
