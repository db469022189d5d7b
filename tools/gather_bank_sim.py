"""Development tool: shared-memory bank model of pass A's column gathers on structured triangle
grids (evidence for the bank steering of the tile-local vertex table; the device
does the same evaluation per tile in k_tile_align).  Two stencils: the uniform 6-neighbour
grid and meshzoo's zigzag grid (`rectangle_tri`: vertices with 8 and 4 neighbours alternate
like a chess board), which is BASELINE config 4.

A tile's table holds its own vertices, then the halo range below, then the one above; lane l
of a warp gathers the column of off-diagonal slot l.  A 16-byte load is served a quarter warp
at a time (8 bank groups of 16 bytes), an 8-byte load half a warp at a time (16 bank pairs);
lanes reading the same word share it.  Prints, for a few tiles, the wavefronts per warp with
the natural layout and with the best even paddings of the two range starts.

On the zigzag grid no placement brings the 16-byte gather to its ideal 4 wavefronts: a quarter
warp that holds an 8-neighbour row needs 8 distinct bank groups for runs of 3 + 2 + 3 entries,
and the two outer runs cannot both avoid the own range's two residues (ncu: 7.8 per warp after
steering, 10.3 before; the 8-byte gather went from 3.9 to its ideal 2.0).

  python tools/gather_bank_sim.py [nx] [margin] [zigzag=1]
"""
import sys


def wavefronts(pos, width, group):
    per = 128 // width
    total = 0
    for g in range(0, len(pos), group):
        count = {}
        for q in set(pos[g:g + group]):
            count[q % per] = count.get(q % per, 0) + 1
        total += max(count.values())
    return total


def stencil(r, nx, zigzag):
    if not zigzag:
        return (r - nx - 1, r - nx, r - 1, r + 1, r + nx, r + nx + 1)
    i, j = divmod(r, nx)
    if (i + j) % 2 == 0:
        return (r - nx - 1, r - nx, r - nx + 1, r - 1, r + 1, r + nx - 1, r + nx, r + nx + 1)
    return (r - nx, r - 1, r + 1, r + nx)


def tile(nx, v0, rows, pad_lo, pad_up, margin, zigzag=False):
    v1 = v0 + rows
    own_start = (v0 - margin) & ~1 if v0 > margin else 0
    own_count = ((v1 + margin - own_start) + 1) & ~1
    cols = [c for r in range(v0, v1) for c in stencil(r, nx, zigzag)]
    below = [c for c in cols if c < own_start]
    above = [c for c in cols if c >= own_start + own_count]
    lo_s, lo_e = (min(below) & ~1) - pad_lo, (max(below) + 2) & ~1  # even starts / counts
    up_s, up_e = (min(above) & ~1) - pad_up, (max(above) + 2) & ~1
    base_lo = own_count
    base_up = base_lo + (lo_e - lo_s)

    def xl(c):
        if own_start <= c < own_start + own_count:
            return c - own_start
        if lo_s <= c < lo_e:
            return base_lo + c - lo_s
        assert up_s <= c < up_e
        return base_up + c - up_s

    return [xl(c) for r in range(v0, v1) for c in stencil(r, nx, zigzag)]


def cost(nx, v0, rows, pads, margin, zigzag=False):
    pos = tile(nx, v0, rows, *pads, margin, zigzag)
    w16 = w8 = n = 0
    for k in range(0, len(pos) - 31, 32):
        w16 += wavefronts(pos[k:k + 32], 16, 8)
        w8 += wavefronts(pos[k:k + 32], 8, 16)
        n += 1
    return w16 / n, w8 / n


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 4001
    margin = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    zz = bool(int(sys.argv[3])) if len(sys.argv) > 3 else True
    for v0, rows in ((10 * nx + 7, 103), (10 * nx + 110, 104), (20 * nx + 1234, 103)):
        cur = cost(nx, v0, rows, (0, 0), margin, zz)
        best = min(
            ((cost(nx, v0, rows, (a, b), margin, zz), (a, b)) for a in range(0, 16, 2) for b in range(0, 16, 2)),
            key=lambda t: (t[0][0] + t[0][1], t[1]),
        )
        print(
            f"nx={nx} margin={margin} {'zigzag' if zz else 'uniform'} first row {v0}, {rows} rows: natural layout x {cur[0]:.2f} / u {cur[1]:.2f} "
            f"wavefronts per warp (ideal 4 / 2); best even paddings {best[1]}: x {best[0][0]:.2f} / u {best[0][1]:.2f}"
        )


if __name__ == "__main__":
    main()
