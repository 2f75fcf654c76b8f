"""How many grad_value RED rows could a warp-level pre-reduction merge?  (VERDICT r01 "next" 2a.)  Host-side count on the
bench distribution: for every (query, head, level) the 16 corner rows of its 4 points, and the rows of the SAME (head, level,
point) sample of the next query (what a warp walking consecutive queries with a running register accumulator could merge)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cape_b200

for dist in ("encoder", "uniform"):
    inp = cape_b200.synthetic.make_inputs(1, 5440, dist=dist, seed=0)
    loc = inp["sampling_locations"][0].numpy()                     # (Lq, M, L, P, 2)
    shapes = inp["spatial_shapes"].numpy()
    print(f"distribution: {dist}")
    tot_rows = tot_within = tot_next = 0
    for l, (H, W) in enumerate(shapes):
        x = loc[:, :, l, :, 0] * W - 0.5
        y = loc[:, :, l, :, 1] * H - 0.5
        x0, y0 = np.floor(x).astype(np.int64), np.floor(y).astype(np.int64)
        rows = []                                                    # (Lq, M, P, 4) linear row index or -1
        for dy in (0, 1):
            for dx in (0, 1):
                xi, yi = x0 + dx, y0 + dy
                ok = (xi >= 0) & (xi < W) & (yi >= 0) & (yi < H)
                rows.append(np.where(ok, yi * W + xi, -1))
        rows = np.stack(rows, -1)
        lq, m, p, _ = rows.shape
        flat = rows.reshape(lq, m, p * 4)
        inb = flat >= 0
        n_rows = int(inb.sum())
        # (i) duplicates among the 16 corners of one (q, m, l)
        srt = np.sort(flat, -1)
        dup = (srt[..., 1:] == srt[..., :-1]) & (srt[..., 1:] >= 0)
        within = int(dup.sum())
        # (ii) rows shared with the same (m, l, p) sample of the NEXT query
        cur, nxt = rows[:-1], rows[1:]                               # (Lq-1, M, P, 4)
        shared = 0
        for c in range(4):
            hit = (cur[..., c:c + 1] == nxt) & (cur[..., c:c + 1] >= 0)
            shared += int(hit.any(-1).sum())
        print(f"  level {l} ({H}x{W}): in-bounds rows {n_rows / (lq * m):5.2f} per (q,m);  mergeable inside a (q,m,l): "
              f"{100 * within / n_rows:4.1f} %;  shared with the next query's same sample: {100 * shared / n_rows:4.1f} %")
        tot_rows += n_rows
        tot_within += within
        tot_next += shared
    print(f"  all levels: {tot_rows / (5440 * 8):5.2f} rows per (q,m); inside-(q,m,l) merges {100 * tot_within / tot_rows:4.1f} %; "
          f"next-query merges {100 * tot_next / tot_rows:4.1f} %")
