"""Development: count what the backward's marching scatter sends to L2 for the bench map, per scheme, on the CPU.

Emulates the warp-level hand-over rules of csrc/warp_bwd_tma.cu on the synthetic 1080p map (tests/synth.py) and counts,
per 32-pixel output row of a warp: RED instructions, 32-byte sectors touched by them, straggler queue entries.
No arithmetic is done -- only which source pixels each lane writes.  Used to compare hand-over schemes before spending
GPU time on them (the kernel's limiter is the SM -> L2 request path: profiles/r02_*.txt)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import synth

H, W = 1080, 1920
TW, TH, SR = 64, 16, 4     # tile, strip rows


def taps(seed=1):
    g = synth.make_map("smooth", 1, H, W, False, seed=seed)[0]
    ix = ((g[..., 0].astype(np.float32) + 1) * W - 1) * 0.5
    iy = ((g[..., 1].astype(np.float32) + 1) * H - 1) * 0.5
    return np.floor(ix).astype(np.int64), np.floor(iy).astype(np.int64)


def sectors(offs):
    """distinct 32-byte sectors (8 floats) among linear float offsets"""
    return len(np.unique(np.asarray(offs) >> 3))


def simulate(x0, y0, ecarry, vdup, rows=slice(128, 128 + 256), strip_rows=SR):
    """interior strips only (the sample rows are away from the frame border)"""
    o_all = y0 * W + x0
    n_rows = 0
    red_instr = 0          # per channel
    red_sectors = 0
    entries = 0
    tops_sectors = 0
    carry_sectors = 0
    r0 = rows.start
    for sr in range(rows.start, rows.stop, strip_rows):
        for c0 in range(0, W - 31, 32):
            co = np.full(32, -1, np.int64); eo = np.full(32, -1, np.int64)
            for r in range(sr, sr + strip_rows):
                o = o_all[r, c0:c0 + 32]
                take = np.zeros(32, bool); take[1:] = o[:-1] + 1 == o[1:]
                given = np.zeros(32, bool); given[:-1] = take[1:]
                chain = co == o
                vd = (co == o + W) if vdup else np.zeros(32, bool)
                broke = (co >= 0) & ~chain & ~vd
                n_e = 0
                if ecarry:
                    e_chain = eo == o + 1
                    e_vd = (eo == o + W + 1) if vdup else np.zeros(32, bool)
                    e_broke = (eo >= 0) & ~e_chain & ~e_vd
                    n_e += (~given).sum() + broke.sum() + e_broke.sum()
                    eo = np.where(~given, o + W + 1, -1)
                else:
                    n_e += 2 * (~given).sum() + broke.sum()
                entries += n_e
                s = sectors(o)
                tops_sectors += s
                red_instr += 1
                co = o + W
                n_rows += 1
            # strip end: parked sums
            carry_sectors += sectors(co)
            red_instr += 1
            if ecarry:
                entries += (eo >= 0).sum()
    q_instr = entries / 32.0 + (n_rows / strip_rows) * 0.5   # dense drains + about half a partial flush per strip
    return dict(rows=n_rows, tops_sectors=tops_sectors / n_rows, carry_sectors=carry_sectors / n_rows,
                entries=entries / n_rows, red_instr=(red_instr + q_instr) / n_rows)


if __name__ == "__main__":
    x0, y0 = taps()
    for ec, vd in ((0, 0), (1, 0), (0, 1), (1, 1)):
        r = simulate(x0, y0, ec, vd)
        total = r["tops_sectors"] + r["carry_sectors"] + r["entries"]
        print(f"ecarry {ec} vdup {vd}: per row and channel: tops {r['tops_sectors']:.2f} sectors, parked {r['carry_sectors']:.2f}, "
              f"straggler entries (= sectors) {r['entries']:.2f}, total {total:.2f} sectors, {r['red_instr']:.2f} RED instr")
