"""Prototype (development aid) of the direct DistributeOctTree construction used by csrc/og_octree2.cuh: the full passes of the
reference (ORBextractor.cc:590-665) produce, at depth D, exactly the non-empty cells of a fixed quadtree, and the list order
is a lexicographic order of the cell paths with alternating directions.  Checked against the oracle on random cases.
    python tools/dev/octree_direct_proto.py [cases]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol  # noqa: E402


def paths(xs, ys, W, H, Dmax):
    """root index and quadrant per depth 1..Dmax for every key"""
    n_ini = int(np.floor(np.float32(W) / np.float32(H) + np.float32(0.5)))   # C round(): halves away from zero
    hX = np.float32(W) / np.float32(n_ini)
    M = len(xs)
    root = (xs.astype(np.float32) / hX).astype(np.int32)
    x0 = (hX * root.astype(np.float32)).astype(np.int32)
    x1 = (hX * (root + 1).astype(np.float32)).astype(np.int32)
    y0 = np.zeros(M, np.int32)
    y1 = np.full(M, H, np.int32)
    q = np.zeros((Dmax + 1, M), np.int32)
    for d in range(1, Dmax + 1):
        sx = x0 + ((x1 - x0 + 1) >> 1)
        sy = y0 + ((y1 - y0 + 1) >> 1)
        rx = xs >= sx
        ry = ys >= sy
        q[d] = rx + 2 * ry
        x0 = np.where(rx, sx, x0)
        x1 = np.where(rx, x1, sx)
        y0 = np.where(ry, sy, y0)
        y1 = np.where(ry, y1, sy)
    return n_ini, root, q


def octree_direct(xs, ys, resp, W, H, N, Dh=6):
    M = len(xs)
    if M == 0:
        return []
    n_ini, root, q = paths(xs, ys, W, H, Dh)
    # cell index at depth d: root * 4^d + path
    cell = [root.copy()]
    for d in range(1, Dh + 1):
        cell.append(cell[-1] * 4 + q[d])
    cnt = [np.bincount(cell[d], minlength=n_ini * 4 ** d) for d in range(Dh + 1)]
    n_d = [int((c > 0).sum()) for c in cnt]
    # children created at depth d with more than one key (= all depth-d cells with > 1 key)
    nexp_d = [int((c > 1).sum()) for c in cnt]
    # ---- full passes -----------------------------------------------------------------------------------------
    D, careful = None, False
    for d in range(1, Dh + 1):
        if n_d[d] >= N or n_d[d] == n_d[d - 1]:
            D = d
            break
        if n_d[d] + 3 * nexp_d[d] > N:
            D, careful = d, True
            break
    if D is None:
        return None   # deeper than the histogram: fall back

    def flip(d):   # xor mask / root direction of rho_d
        m = 0
        for j in range(1, d + 1):
            m = (m << 2) | (3 if (d - j) % 2 == 0 else 0)
        return m, (d % 2 == 1)   # the roots are pushed back (not front): the root component has the direction of q1

    # list at depth D: nodes as (depth, cell)
    lst = []
    for d in range(D, -1, -1):
        ncell = n_ini * 4 ** d
        m, rdesc = flip(d)
        for f in range(ncell):   # flipped index ascending
            r, p = divmod(f, 4 ** d)
            c = ((n_ini - 1 - r) if rdesc else r) * 4 ** d + (p ^ m)
            if d == 0:
                ok = cnt[0][c] == 1 if D > 0 else cnt[0][c] > 0
            else:
                born = cnt[d][c] > 0 and cnt[d - 1][c >> 2] > 1
                ok = born if d == D else (born and cnt[d][c] == 1)
            if ok:
                lst.append((d, c))
    assert len(lst) == n_d[D], (len(lst), n_d[D])
    # ---- careful rounds --------------------------------------------------------------------------------------
    if careful:
        # candidates in creation order = reverse list order of the depth-D multi nodes
        cand = [nd for nd in reversed(lst) if nd[0] == D and cnt[D][nd[1]] > 1]
        depth = D
        while True:
            prev = len(lst)
            if depth + 1 > Dh:
                return None
            order = sorted(range(len(cand)), key=lambda j: (cnt[depth][cand[j][1]], j), reverse=True)
            size = len(lst)
            front = []     # pushed to the front, in push order
            newcand = []
            dead = set()
            for j in order:
                c = cand[j][1]
                for k in range(4):
                    cc = c * 4 + k
                    if cnt[depth + 1][cc] > 0:
                        front.append((depth + 1, cc))
                        if cnt[depth + 1][cc] > 1:
                            newcand.append((depth + 1, cc))
                dead.add(cand[j])
                size = len(lst) - len(dead) + len(front)
                if size >= N:
                    break
            lst = list(reversed(front)) + [nd for nd in lst if nd not in dead]
            if len(lst) >= N or len(lst) == prev:
                break
            cand = newcand
            depth += 1
    # ---- best key per node -----------------------------------------------------------------------------------
    out = []
    for d, c in lst:
        idx = np.nonzero(cell[d] == c)[0]
        b = idx[np.argmax(resp[idx])]   # first maximum in emission order
        out.append((int(xs[b]), int(ys[b]), int(resp[b])))
    return out


def main():
    cases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    oracle = ol.load_port()
    rs = np.random.RandomState(5)
    fallbacks = 0
    for it in range(cases):
        width, height = int(rs.randint(30, 1300)), int(rs.randint(30, 400))
        if round(width / height) < 1:
            continue
        M = int(rs.randint(0, 3000))
        N = int(rs.randint(1, 500))
        if it % 3 == 0:
            xs = np.clip(rs.normal(width / 2, width / 12, M * 2), 3, width - 4).astype(int)
            ys = np.clip(rs.normal(height / 2, height / 12, M * 2), 3, height - 4).astype(int)
        else:
            xs = rs.randint(3, width - 3, M * 2)
            ys = rs.randint(3, height - 3, M * 2)
        pts = np.unique(np.stack([ys, xs], 1), axis=0)
        pts = pts[rs.permutation(len(pts))][:M]
        cand = np.zeros(len(pts), ol.KP_DTYPE)
        cand["x"], cand["y"] = pts[:, 1], pts[:, 0]
        cand["response"] = rs.randint(7, 60 if it % 2 else 255, len(pts))
        exp = ol.octree(oracle, "orbo", cand, 16, 16 + width, 16, 16 + height, N)
        got = octree_direct(pts[:, 1].astype(np.int32), pts[:, 0].astype(np.int32), cand["response"].astype(np.int32), width, height, N)
        if got is None:
            fallbacks += 1
            continue
        e = [(int(a), int(b), int(c)) for a, b, c in zip(exp["x"], exp["y"], exp["response"])]
        assert got == e, (it, width, height, len(pts), N, got[:5], e[:5], len(got), len(e))
    print("ok", cases, "cases,", fallbacks, "fallbacks")


if __name__ == "__main__":
    main()
