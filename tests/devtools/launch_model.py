"""Development aid (NOT product, NOT test): discrete-event model of ONE walk launch.  One-warp CTAs are handed out in
blockIdx order to 148 SMs x 4 schedulers x R/4 resident slots; the warps of a scheduler share its issue port
(processor sharing) and a lone warp progresses at most at rate rho of the port.  Per-group work comes from the
schedule replay of walk_model.c over the oracle's tree (35.7 issue cycles per list entry, 125 per batch, ...).

    python tests/devtools/launch_model.py [N]        # N = 125000: the grid of a 1/8 shard of N = 1M

Measured on one B200 (profiles/README.md): walk ms = 2.38e-6 x (CTAs x flops per particle) + 0.2 ms for every grid
from 3907 to 31251 CTAs.  The model reproduces a constant of that size with rho ~ 0.25 (0.16-0.19 ms): a launch ends
with whatever its last CTAs are, groups differ by up to 2x in work, and the last warps of a scheduler run far below
the port's rate.  Handing out the heaviest groups first (previous step's work as the estimate) removes most of it in
the model: -6 % at 31251 CTAs, -15 % at 3907."""
import ctypes as C, os, sys, subprocess, heapq
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', '..'))
from oracle import okd
here = os.path.dirname(os.path.abspath(__file__))
so = '/tmp/walk_model.so'
subprocess.run(['gcc', '-O2', '-shared', '-fPIC', '-o', so, os.path.join(here, 'walk_model.c'), '-lm'], check=True)
lib = C.CDLL(so)
names = "batches popped far near mixed leaves entries lanes mixed_rounds mixed_round_nodes leaf_rounds leaf_round_nodes max_sp max_pend drains part_entries tests pack2 pack4 pack32".split() + [f"h{i}" for i in range(33)] + [f"tl{i}" for i in range(24)] + [f"sw{i}" for i in range(16)]

def group_costs(N):
    o = okd.Oracle()
    parts = o.circular_orbits(N)
    nodes, idx, last = o.build_tree_canonical(parts, threads=8)
    n = len(parts)
    ng = (n + 31) // 32
    cost = np.zeros(ng)
    st = (C.c_double * len(names))()
    for g in range(ng):
        lib.wm_run(C.c_void_p(nodes.ctypes.data), C.c_void_p(parts.ctypes.data), C.c_void_p(idx.ctypes.data), C.c_uint64(n),
                   C.c_double(0.3), C.c_uint64(g), C.c_uint64(1), C.c_uint64(1), C.c_int(0), C.c_int(320), st)
        d = dict(zip(names, st))
        # issue cycles: drain 35.7 per entry; traversal ~ (batch overhead 125 + 5.5 per popped node) ; mixed 33 per mixed node; leaves 50 per leaf
        cost[g] = 35.7 * d['entries'] + 125 * d['batches'] + 5.5 * d['popped'] + 33 * d['mixed'] + 50 * d['leaves']
    return cost

def simulate(cost, R=24, sms=148, rho=1.0, order=None):
    """returns kernel cycles.  Each scheduler: set of resident warps with remaining work; rate per warp = min(1/k, rho)."""
    nsch = sms * 4
    per = R // 4
    if order is None:
        order = np.arange(len(cost))
    nxt = 0
    resident = [[] for _ in range(nsch)]   # remaining work per warp
    t = 0.0
    # initial fill: CTA i -> SM i % sms round-robin, scheduler within SM by fill
    for s_round in range(per):
        for sch_in_sm in range(4):
            for sm in range(sms):
                if nxt < len(order):
                    resident[sm * 4 + sch_in_sm].append(cost[order[nxt]]); nxt += 1
    # event loop: per scheduler independent until it needs a new CTA: global order matters only via dispatch order.
    # next finishing time per scheduler
    def next_finish(sc):
        ws = resident[sc]
        if not ws: return None
        k = len(ws); rate = min(1.0 / k, rho)
        return min(ws) / rate
    last = np.zeros(nsch)   # local clock of each scheduler
    heap = []
    for sc in range(nsch):
        nf = next_finish(sc)
        if nf is not None: heapq.heappush(heap, (nf, sc))
    end = 0.0
    while heap:
        tf, sc = heapq.heappop(heap)
        ws = resident[sc]
        k = len(ws); rate = min(1.0 / k, rho)
        dt = tf - last[sc]
        adv = dt * rate
        ws = [w - adv for w in ws]
        # remove finished (the min)
        j = int(np.argmin(ws)); ws.pop(j)
        if nxt < len(order):
            ws.append(cost[order[nxt]]); nxt += 1
        resident[sc] = ws
        last[sc] = tf
        end = max(end, tf)
        if ws:
            k = len(ws); rate = min(1.0 / k, rho)
            heapq.heappush(heap, (tf + min(ws) / rate, sc))
    return end

if __name__ == '__main__':
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
    f = f'/tmp/gcost_{N}.npy'
    if os.path.exists(f): cost = np.load(f)
    else:
        cost = group_costs(N); np.save(f, cost)
    print(f"N={N} groups={len(cost)} mean {cost.mean():.0f} cycles  std/mean {cost.std()/cost.mean():.3f}  max/mean {cost.max()/cost.mean():.2f}")
    ideal = cost.sum() / (148 * 4)
    print(f" issue-bound ideal {ideal/1.965e6:.4f} ms")
    for rho in (1.0, 0.5, 0.35, 0.25):
        t24 = simulate(cost, 24, rho=rho); t32 = simulate(cost, 32, rho=rho)
        tl = simulate(cost, 24, rho=rho, order=np.argsort(-cost))
        print(f" rho={rho}: R=24 {t24/1.965e6:.4f} ms   R=32 {t32*1.05/1.965e6:.4f} ms (x1.05 per-CTA)   longest-first R=24 {tl/1.965e6:.4f} ms")
