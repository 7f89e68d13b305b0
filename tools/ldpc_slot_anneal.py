#!/usr/bin/env python3
"""Row-slot order of the node-centred BP kernel (csrc/decode.cu, kRowSlotOf) chosen by simulated annealing over the shared-memory
bank conflicts of its two scatter phases.  Per iteration a warp issues 18 stores of variable->check messages (lanes = variables
lane + 32 r, targets toc[row slot][position]) and 19 stores of check->variable messages (lanes = row slots, targets tov[n][q]); which
BANK a target falls on depends on the row's slot number (4 floats per slot: slots congruent mod 8 share banks).  The natural order
(7-variable rows first, ascending) costs 132 wavefronts for those 37 stores; the order printed here 94.  Constraint kept: the
7-variable rows stay in the first 32 slots (the 7-wide first round).  Pure table change: no arithmetic depends on it.
usage: tools/ldpc_slot_anneal.py [seed] [iterations]"""
import os, random, re, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(ROOT, "rtlsdr-ft8d_b200", "csrc", "ft8_tables.h")).read()


def arr(name):
    m = re.search(name + r"\s*\[[^\]]*\](?:\[[^\]]*\])?\s*=\s*\{(.*?)\};", src, re.S)
    return [int(x) for x in re.findall(r"-?\d+", m.group(1))]


Nm = np.array(arr("kFt8tNm")).reshape(83, 7); Mn = np.array(arr("kFt8tMn")).reshape(174, 3); NR = np.array(arr("kFt8tNumRows"))
kTocLo = 176 * 4; kTocHi = kTocLo + 84 * 4
pos_of = {(n, q): (Mn[n][q] - 1, list(Nm[Mn[n][q] - 1][:NR[Mn[n][q] - 1]] - 1).index(n)) for n in range(174) for q in range(3)}


def cost(slot_of):
    tot = 0
    for r in range(6):
        for q in range(3):
            cnt = {}
            for lane in range(32):
                n = lane + 32 * r
                if n < 174:
                    m, pos = pos_of[(n, q)]; sl = slot_of[m]
                    a = (kTocLo + 4 * sl + pos) if pos < 4 else (kTocHi + 4 * sl + pos - 4)
                    cnt.setdefault(a % 32, set()).add(a)
            tot += max(len(v) for v in cnt.values())
    inv = {slot_of[m]: m for m in range(83)}
    for base, kpos in ((0, 7), (32, 6), (64, 6)):
        for j in range(kpos):
            cnt = {}
            for lane in range(32):
                sl = base + lane
                if sl in inv:
                    m = inv[sl]
                    if j < NR[m]:
                        n = Nm[m][j] - 1; a = 4 * n + list(Mn[n] - 1).index(m)
                    else:
                        a = kTocHi + 4 * sl + 3
                    cnt.setdefault(a % 32, set()).add(a)
            tot += max(len(v) for v in cnt.values())
    return tot


if __name__ == "__main__":
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60000
    p = np.zeros(83, int); k = 0
    for want in (7, 6):
        for m in range(83):
            if NR[m] == want:
                p[m] = k; k += 1
    c = cost(p); print("natural order:", c, "wavefronts (37 stores)")
    random.seed(seed); best, bestp, T = c, p.copy(), 2.0
    for it in range(iters):
        a, b = random.sample(range(83), 2)
        q = p.copy(); q[a], q[b] = q[b], q[a]
        if any(NR[m] == 7 and q[m] >= 32 for m in (a, b)):
            continue
        cq = cost(q)
        if cq <= c or random.random() < np.exp((c - cq) / T):
            p, c = q, cq
            if c < best:
                best, bestp = c, p.copy()
        T = max(0.05, T * 0.9999)
    print("best:", best)
    print("constexpr uint8_t kRowSlotOf[83] = {" + ", ".join(str(int(x)) for x in bestp) + "};")
