"""A third statement of the Kuhn subgame step, written the way the DEVICE path works (csrc/mccfr.cu + rbp_subgame_*), in numpy float32:
dense per-world tables seeded by `subgame_seed_kernel`'s rule, the `visits == 0 -> blueprint weight` read in the sampling preamble, the
LIFO tree of `mccfr_sample_kernel` started at the entry node, the ordered fold.  tests/test_oracle_subgame.py checks it bit for bit against
oracle/subgame.hpp (which follows the reference's WorldProfile / HashMap semantics): that the dense-table design is equivalent to the
reference's two-level profile is the one non-mechanical claim of the device path, and this is its CPU check.  Test infrastructure only.
"""
import numpy as np

F = np.float32
EPS = np.finfo(np.float32).tiny
OPEN, CHECK, BET, CHECKBET, FOLD0, FOLD1, SHOW1, SHOW2 = range(8)
TURN = {OPEN: 0, CHECKBET: 0, CHECK: 1, BET: 1}
HIST = {OPEN: 0, CHECK: 1, BET: 2, CHECKBET: 3}
CHILD = {OPEN: (CHECK, BET), CHECK: (SHOW1, CHECKBET), BET: (FOLD1, SHOW2), CHECKBET: (FOLD0, SHOW2)}   # choices() order: [Check, Bet] / [Fold, Call]


def info_key(node, hole):
    return 1 | (HIST[node] << 1) | ((hole[TURN[node]] >> 1) << 3)


def payoff(node, hole, p):
    if node in (FOLD0, FOLD1):
        who = 0 if node == FOLD0 else 1
        return F(-1.0) if who == p else F(1.0)
    stake = F(2.0) if node == SHOW2 else F(1.0)
    r0, r1 = hole[0] >> 1, hole[1] >> 1
    if r0 > r1:
        return stake if p == 0 else F(-stake)
    if r0 < r1:
        return stake if p == 1 else F(-stake)
    return F(0.0)


def unit(r):
    return F(r >> 8) * F(1.0 / 16777216.0)


class DenseSubgame:
    def __init__(self, philox, blueprint_rows, worlds, weights, entries, entry_node, seed, k=16384.0, temperature=1.0, smoothing=2.0, curiosity=0.05):
        self.philox, self.worlds, self.weights = philox, worlds, [F(w) for w in weights]
        self.entries, self.entry_node, self.seed, self.t = entries, entry_node, seed, 0
        self.tau, self.beta, self.eps_q = F(temperature), F(smoothing), F(curiosity)
        bp = {(int(r["info_key"]), int(r["action"])): r for r in blueprint_rows}
        keys = sorted({1 | (h << 1) | (r << 3) for h in range(4) for r in range(3)})
        self.fb, seed_rows = {}, {}
        k = F(k)
        for key in keys:                                             # subgame_seed_kernel, one infoset
            w = [max(F(bp[(key, a)]["weight"]) if (key, a) in bp else F(0.0), EPS) for a in range(2)]
            s = F(F(0.0) + w[0]) + w[1]
            for a in range(2):
                policy = F(w[a] / s)
                reg = F(bp[(key, a)]["regret"]) if (key, a) in bp else F(0.0)
                seed_rows[(key, a)] = [F(F(F(policy * k) * F(k + F(1.0))) / F(2.0)), max(reg, EPS), F(0.0), 0]   # weight, regret, payoff, visits
                self.fb[(key, a)] = w[a]
        self.table = [{ka: list(v) for ka, v in seed_rows.items()} for _ in range(worlds)]
        self.drawn = [0] * worlds

    def views(self, tab, key):
        r = [max(tab[(key, a)][1], EPS) for a in range(2)]
        rd = F(F(0.0) + r[0]) + r[1]
        w = [max(self.fb[(key, a)] if tab[(key, a)][3] == 0 else tab[(key, a)][0], EPS) for a in range(2)]
        ws = F(F(0.0) + w[0]) + w[1]
        denom = F(ws + self.beta)
        sw = [max(F(F(F(w[a] / self.tau) + self.beta) / denom), self.eps_q) for a in range(2)]
        z = F(F(0.0) + sw[0]) + sw[1]
        return [F(r[a] / rd) for a in range(2)], [F(sw[a] / z) for a in range(2)]

    def step(self):
        t = self.t
        p = self.philox([t & 0xFFFFFFFF, 0, 0xFFFFFFFE, 5], [self.seed & 0xFFFFFFFF, self.seed >> 32])
        total = F(0.0)
        for w in self.weights:
            total = F(total + w)
        x, cum, world = F(unit(p[0]) * total), F(0.0), self.worlds - 1
        for i, w in enumerate(self.weights):
            cum = F(cum + w)
            if x < cum:
                world = i
                break
        self.drawn[world] += 1
        tab, hole, walker = self.table[world], self.entries[world], t % 2
        pol = {}

        def view(node):
            key = info_key(node, hole)
            if key not in pol:
                pol[key] = self.views(tab, key)
            return pol[key]

        # LIFO tree (mccfr_sample_kernel): local arrays, head = newest child
        nodes, parent, act, head, nxt, todo = [self.entry_node], [-1], [0], [-1], [-1], []

        def expand(k):
            nd = nodes[k]
            if nd not in CHILD:
                return
            if TURN[nd] == walker:
                for a in range(2):
                    todo.append((CHILD[nd][a], k, a))
                return
            _, q = view(nd)
            r = self.philox([t & 0xFFFFFFFF, world, info_key(nd, hole), 0], [self.seed & 0xFFFFFFFF, self.seed >> 32])
            tot = F(0.0)
            for a in range(2):
                tot = F(tot + max(q[a], EPS))
            xx, c, pick = F(unit(r[0]) * tot), F(0.0), 1
            for a in range(2):
                c = F(c + max(q[a], EPS))
                if xx < c:
                    pick = a
                    break
            todo.append((CHILD[nd][pick], k, pick))

        expand(0)
        while todo:
            nd, par, a = todo.pop()
            k = len(nodes)
            nodes.append(nd); parent.append(par); act.append(a); head.append(-1); nxt.append(head[par]); head[par] = k
            expand(k)

        def value(n, rel, smp, hero):                                  # recursed_value
            if head[n] < 0:
                return F(F(rel / smp) * payoff(nodes[n], hole, hero))
            sg, q = view(nodes[n])
            acc, c = F(0.0), head[n]
            while c >= 0:
                r2 = F(rel * sg[act[c]])
                s2 = F(smp * q[act[c]]) if TURN[nodes[n]] != hero else smp
                acc = F(acc + value(c, r2, s2, hero))
                c = nxt[c]
            return acc

        records = []
        for k in range(len(nodes)):
            nd = nodes[k]
            if nd not in CHILD or TURN[nd] != walker or head[k] < 0:
                continue
            cf, sm, n = F(1.0), F(1.0), k
            while parent[n] >= 0:
                pn = nodes[parent[n]]
                if TURN[pn] != walker:
                    sg, q = view(pn)
                    cf, sm = F(cf * sg[act[n]]), F(sm * q[act[n]])
                n = parent[n]
            reach = F(cf / sm)
            vals, acts, c = [], [], head[k]
            while c >= 0:
                vals.append(F(reach * value(c, F(1.0), F(1.0), walker))); acts.append(act[c])
                c = nxt[c]
            sg, _ = view(nd)
            ev = F(0.0)
            for v, a in zip(vals, acts):
                ev = F(ev + F(sg[a] * v))
            dr = [F(0.0), F(0.0)]
            for v, a in zip(vals, acts):
                dr[a] = F(dr[a] + F(v - ev))
            records.append((info_key(nd, hole), sg, F(F(0.0) + ev), dr))
        tf = F(t)
        for key, sg, pay, dr in records:                               # ordered fold: SummedRegret, LinearWeight, Welford payoff, visits
            for a in range(2):
                row = tab[(key, a)]
                row[1] = F(row[1] + dr[a])
                row[0] = max(F(row[0] + F(sg[a] * tf)), EPS)
                row[2] = F(row[2] + F(F(pay - row[2]) / F(row[3] + 1)))
                row[3] += 1
        self.t += 1

    def rows(self, world):
        out = [(key, a, v[0], v[1], v[2], v[3]) for (key, a), v in self.table[world].items() if v[3] > 0]
        return sorted(out)
