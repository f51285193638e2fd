"""Dev tool: throughput with P pairs in flight per GPU (P host threads, one libnct context + stream each)."""
import os, sys, json, threading, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
from oracle import synth
pkg = g.load_package(); dev = torch.device("cuda:0")
w = synth.vgg19_weights(19)
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
for P in [int(x) for x in sys.argv[1].split(",")]:
    ctxs, pairs = [], []
    for i in range(P):
        c = pkg.Context(0); c.load_vgg19_weights(w); c.set_vgg_engine(2); ctxs.append(c)
        a, b = synth.pair(i, 700, 700)
        pairs.append((torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)))
    def work(i, n):
        for _ in range(n):
            ctxs[i].transfer_pair_dev(*pairs[i])
        ctxs[i].synchronize()
    ts = [threading.Thread(target=work, args=(i, 1)) for i in range(P)]
    [t.start() for t in ts]; [t.join() for t in ts]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ts = [threading.Thread(target=work, args=(i, K)) for i in range(P)]
    [t.start() for t in ts]; [t.join() for t in ts]
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps(dict(pairs_in_flight=P, pairs=P * K, seconds=round(dt, 3), ms_per_pair=round(1e3 * dt / (P * K), 1),
                          mp_per_s=round(P * K * 0.49 / dt, 3))), flush=True)
    for c in ctxs: c.close()
