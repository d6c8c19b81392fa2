"""Generates the min/max straight-line code of k_median5's shared-work selection (flow_kernels.cu).

A thread produces the 5x5 medians of four vertically adjacent pixels from eight rows r=0..7 of five
samples each (window of output j: rows j..j+4):
  S[r]  = the row sorted (9 compare-exchanges)
  P[r]  = merge(S[r], S[r+1]) for r = 1, 3, 5 (sorted 10)
  Q[r]  = ranks 8..13 of merge(P[r], P[r+2]) for r = 1, 3 (the six middle values of 20: seven of the
          twenty have at least 13 values above them in any 25-superset and seven have 13 below, so
          the median of the 25 is the 6th smallest of Q u S)
  out0 = sel6(Q[1], S[0]); out1 = sel6(Q[1], S[5]); out2 = sel6(Q[3], S[2]); out3 = sel6(Q[3], S[7])
  sel6(A sorted 6, B sorted 5) = 6th smallest of the union = min over i+j=6 of max(A_i, B_j)
The merge networks are Batcher odd-even merges, padded with +inf, pruned by liveness, dead halves of
compare-exchanges dropped and chains min(min(a,b),c) fused into the 3-input FMNMX3 (PTX min.f32 a,b,c).
Every piece is verified here: merges on all sorted 0/1 inputs (0/1 principle) and the whole run against
sorted()[12] on random and tie-heavy inputs.  Usage: python tools/gen_median_run.py [R] > csrc/median_run.inc   (R medians per thread, default 4; 8 -> median_run8.inc)
"""
import itertools, random, sys

INF = 'INF'

def oddeven_merge(lo, hi, r):
    step = r * 2
    if step < hi - lo:
        yield from oddeven_merge(lo, hi, step)
        yield from oddeven_merge(lo + r, hi, step)
        for i in range(lo + r, hi - r, step):
            yield (i, i + r)
    else:
        yield (lo, lo + r)

def merge_network(n_a, n_b, need):
    """comparators (i,j) on wires; A on wires 0..n_a-1, B on wires n..n+n_b-1 of a 2n-wire Batcher merge
    (n = power of two >= max). Returns (ops, out_wires): straight-line ops over value names."""
    n = 1
    while n < max(n_a, n_b): n *= 2
    comps = list(oddeven_merge(0, 2 * n - 1, 1))
    return n, comps

class Builder:
    def __init__(self):
        self.ops = []      # (dst, kind, a, b)
        self.cnt = 0
    def new(self):
        self.cnt += 1
        return f't{self.cnt}'
    def ce(self, a, b):
        lo, hi = self.new(), self.new()
        self.ops.append((lo, 'min', a, b))
        self.ops.append((hi, 'max', a, b))
        return lo, hi

def sort5(bd, v):
    v = list(v)
    for i, j in [(0,1),(3,4),(2,4),(2,3),(0,3),(0,2),(1,4),(1,3),(1,2)]:
        v[i], v[j] = bd.ce(v[i], v[j])
    return v

def merge_sorted(bd, A, B):
    """Batcher odd-even merge of two sorted lists, +inf padding resolved statically. Returns sorted list of names."""
    n, comps = merge_network(len(A), len(B), None)
    wires = list(A) + [INF] * (n - len(A)) + list(B) + [INF] * (n - len(B))
    for i, j in comps:
        a, b = wires[i], wires[j]
        if b is INF:
            continue
        if a is INF:
            wires[i], wires[j] = b, a
            continue
        wires[i], wires[j] = bd.ce(a, b)
    out = [w for w in wires if w is not INF]
    assert all(w is INF for w in wires[len(A) + len(B):]), 'pads must end up on top'
    return wires[:len(A) + len(B)]

def sel6(bd, A6, B5):
    terms = [A6[5]]
    for i in range(5, 0, -1):            # (i, j) with i + j = 6, j = 1..5
        j = 6 - i
        t = bd.new(); bd.ops.append((t, 'max', A6[i - 1], B5[j - 1])); terms.append(t)
    # min of six terms
    t1 = bd.new(); bd.ops.append((t1, 'min', terms[0], terms[1]))
    t2 = bd.new(); bd.ops.append((t2, 'min', t1, terms[2]))
    t3 = bd.new(); bd.ops.append((t3, 'min', t2, terms[3]))
    t4 = bd.new(); bd.ops.append((t4, 'min', t3, terms[4]))
    t5 = bd.new(); bd.ops.append((t5, 'min', t4, terms[5]))
    return t5

def build(R=4):
    """R (even) vertically adjacent medians from R + 4 rows."""
    bd = Builder()
    rows = [[f'v{r}_{c}' for c in range(5)] for r in range(R + 4)]
    S = [sort5(bd, rows[r]) for r in range(R + 4)]
    P = {r: merge_sorted(bd, S[r], S[r + 1]) for r in range(1, R + 2, 2)}
    Q = {r: merge_sorted(bd, P[r], P[r + 2])[7:13] for r in range(1, R, 2)}
    outs = []
    for j in range(R):   # window of output j: rows j .. j+4
        outs.append(sel6(bd, Q[j + 1], S[j]) if j % 2 == 0 else sel6(bd, Q[j], S[j + 4]))
    return bd.ops, outs


def dce(ops, outs):
    live = set(outs)
    keep = []
    for op in reversed(ops):
        if op[0] in live:
            keep.append(op)
            live.update(op[2:])
    keep.reverse()
    return keep

def fuse3(ops, outs):
    """min(min(a,b),c) with a single-use inner value -> min3(a,b,c)."""
    uses = {}
    for dst, kind, *args in ops:
        for a in args: uses[a] = uses.get(a, 0) + 1
    for o in outs: uses[o] = uses.get(o, 0) + 1
    defs = {op[0]: op for op in ops}
    removed = set()
    out = []
    for op in ops:
        dst, kind, *args = op
        if len(args) == 2:
            for idx in (0, 1):
                inner = defs.get(args[idx])
                if inner and inner[1] == kind and len(inner) == 4 and uses[inner[0]] == 1 and inner[0] not in removed:
                    other = args[1 - idx]
                    op = (dst, kind + '3', inner[2], inner[3], other)
                    removed.add(inner[0])
                    break
        out.append(op)
        defs[op[0]] = op
    return [op for op in out if op[0] not in removed]

def evaluate(ops, outs, env):
    env = dict(env)
    for dst, kind, *args in ops:
        vals = [env[a] for a in args]
        env[dst] = min(vals) if kind.startswith('min') else max(vals)
    return [env[o] for o in outs]

def verify(ops, outs):
    R = len(outs)
    rng = random.Random(7)
    names = [f'v{r}_{c}' for r in range(R + 4) for c in range(5)]
    for trial in range(20000):
        mode = trial % 4
        if mode == 0: vals = [rng.random() for _ in names]
        elif mode == 1: vals = [rng.randint(0, 3) for _ in names]
        elif mode == 2: vals = [rng.randint(0, 1) for _ in names]
        else: vals = [rng.randint(-50, 50) for _ in names]
        env = dict(zip(names, vals))
        got = evaluate(ops, outs, env)
        for j in range(R):
            win = sorted(env[f'v{r}_{c}'] for r in range(j, j + 5) for c in range(5))
            assert got[j] == win[12], (trial, j, got[j], win[12])
    # merges: all sorted 0/1 inputs (0/1 principle for merging networks)
    for na, nb in ((5, 5), (10, 10)):
        for za in range(na + 1):
            for zb in range(nb + 1):
                bd = Builder()
                A = [f'a{i}' for i in range(na)]; B = [f'b{i}' for i in range(nb)]
                M = merge_sorted(bd, A, B)
                env = {f'a{i}': int(i >= za) for i in range(na)}
                env.update({f'b{i}': int(i >= zb) for i in range(nb)})
                assert evaluate(bd.ops, M, env) == sorted(env.values())

def emit(ops, outs):
    print('// generated by tools/gen_median_run.py -- do not edit')
    n2 = sum(1 for op in ops if len(op) == 4); n3 = len(ops) - n2
    print(f'// {len(ops)} min/max instructions for {len(outs)} medians ({n2} two-input, {n3} three-input)')
    for r in range(len(outs) + 4):
        print(' '.join(f'const float v{r}_{c} = v[{r}][{c}];' for c in range(5)))
    for dst, kind, *args in ops:
        if len(args) == 2:
            print(f'const float {dst} = {"fminf" if kind == "min" else "fmaxf"}({args[0]}, {args[1]});')
        else:
            print(f'const float {dst} = {"min3f" if kind == "min3" else "max3f"}({args[0]}, {args[1]}, {args[2]});')
    for j, o in enumerate(outs):
        print(f'out[{j}] = {o};')

if __name__ == '__main__':
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    ops, outs = build(R)
    raw = len(ops)
    ops = dce(ops, outs)
    live = len(ops)
    verify(ops, outs)
    ops = fuse3(ops, outs)
    ops = dce(ops, outs)
    verify(ops, outs)
    print(f'raw {raw} ops, after dead-code elimination {live}, after 3-input fusion {len(ops)}', file=sys.stderr)
    emit(ops, outs)
