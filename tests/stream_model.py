"""CPU model of the rank-counting protocol of csrc/retrieval_tc.cu (stream epilogue, TRB_TC_COUNT_FIRST): NOT the product, a
restatement of the kernel's bookkeeping in numpy so that the tie / shard / padding logic is pinned without a GPU.

One query row at a time.  The gallery is split into contiguous shards; every shard is streamed in 256-row tiles of eight 32-value
chunks, the chunks of a tile are dealt to two "warps" (column groups: chunks 0-3 and 4-7), each warp keeps its own register state
per relevant item r of the row:

    te[r]  the value the stream compares against: nextbelow(thr) until the warp reaches the item (values EQUAL to thr that precede
           the item rank before it), thr itself afterwards;
    sw[r]  first chunk start at which te switches (= local row of the item - 31; -inf for an item of a lower shard, never for an
           item of a higher shard).

Count-first: every chunk is first counted against the CURRENT te (the fast path does nothing else).  The rare path then corrects:
  * a slot that switches in this chunk was counted with ">=" semantics: the values equal to thr AT or AFTER the item go back
    (all of them when the item precedes the chunk);
  * the padded tail of the last tile (zero vectors, similarity exactly 0) goes back where 0 > te.
The result per relevant item must be its 0-based rank under (similarity desc, gallery index asc).
"""
import numpy as np

CH, TILE = 32, 256


def next_below(x: np.float32) -> np.float32:
    return np.nextafter(np.float32(x), np.float32(-np.inf), dtype=np.float32)


def stream_counts(sim_row: np.ndarray, rel_cols, shard_bounds):
    """sim_row [G] float32 similarities of one query; rel_cols: global gallery indices of its relevant items;
    shard_bounds: [(lo, hi), ...] contiguous shards.  Returns the counted rank of every relevant item (summed over shards and
    column groups, like the kernel's atomics + all-reduce)."""
    counts = np.zeros(len(rel_cols), dtype=np.int64)
    thr = [np.float32(sim_row[c]) for c in rel_cols]
    for lo, hi in shard_bounds:
        G = hi - lo
        Gp = -(-G // TILE) * TILE
        local = np.zeros(Gp, dtype=np.float32)          # zero padding rows: similarity exactly 0
        local[:G] = sim_row[lo:hi]
        for cg in range(2):                              # two column groups = two warps with their own state
            te, sw = [], []
            for r, c in enumerate(rel_cols):
                l = c - lo
                te.append(next_below(thr[r]))
                sw.append(-np.inf if l < 0 else (np.inf if l >= G else l - (CH - 1)))
            for t in range(Gp // TILE):
                tail = (t + 1) * TILE > G
                for k in range(4):
                    g0 = t * TILE + (cg * 4 + k) * CH
                    v = local[g0:g0 + CH].copy()
                    # ---- fast part: count against the current compare values (unmasked chunk) ----
                    for r in range(len(rel_cols)):
                        counts[r] += int((v > te[r]).sum())
                    # ---- rare part: corrections ----
                    if tail:
                        npad = min(max(g0 + CH - G, 0), CH)
                        for r in range(len(rel_cols)):
                            if np.float32(0) > te[r]:
                                counts[r] -= npad
                        if g0 + CH > G:
                            v[max(G - g0, 0):] = -np.inf      # padded rows never tie with anything
                    for r, c in enumerate(rel_cols):
                        if g0 >= sw[r]:
                            l = c - lo - g0                # position of the item in this chunk (< 0: before it)
                            j0 = max(l, 0)
                            counts[r] -= int((v[j0:] == thr[r]).sum())
                            te[r] = thr[r]                 # the following chunks compare strictly
                            sw[r] = np.inf
    return counts


def oracle_ranks(sim_row: np.ndarray, rel_cols):
    out = []
    for c in rel_cols:
        s = sim_row[c]
        before = (sim_row > s) | ((sim_row == s) & (np.arange(sim_row.shape[0]) < c))
        out.append(int(before.sum()))
    return np.asarray(out, dtype=np.int64)
