"""Shared-memory bank model of the 3-D half-edge fold (16-byte loads): wavefronts per
quarter-warp access for one lane per slot (today) and two lanes per slot.

  python tools/bank_sim.py [points per side]
"""
import os
import sys

import numpy

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pyfvm_b200 import meshes  # noqa: E402


def slot_counts(n):
    xs = numpy.linspace(0, 1, n)
    m = meshes.Mesh(*meshes.cube_tetra(xs, xs, xs))
    idx = numpy.asarray(m.idx_hierarchy).reshape(2, -1)
    nv = len(m.points)
    key = numpy.concatenate([idx[0] * nv + idx[1], idx[1] * nv + idx[0]])
    _, c = numpy.unique(key, return_counts=True)  # CSR slots in row-major order
    return c


def one_lane_per_slot(unit, pairs):
    tot = acc = 0
    k = len(pairs) // 32 * 32
    U, C = unit[:k].reshape(-1, 32), pairs[:k].reshape(-1, 32)
    for ch in range(len(U)):
        for j in range(C[ch].max()):
            act = C[ch] > j
            for q in range(4):
                a = act[q * 8 : (q + 1) * 8]
                if a.any():
                    tot += numpy.bincount((U[ch, q * 8 : (q + 1) * 8][a] + j) % 8, minlength=8).max()
                    acc += 1
    return tot, acc


def two_lanes_per_slot(unit, pairs):
    tot = acc = 0
    k = len(pairs) // 16 * 16
    U, C = unit[:k].reshape(-1, 16), pairs[:k].reshape(-1, 16)
    for ch in range(len(U)):
        for j in range(((C[ch] + 1) // 2).max()):
            for q in range(4):
                cls = [(U[ch, q * 4 + i] + 2 * j + sub) % 8
                       for i in range(4) for sub in (0, 1) if 2 * j + sub < C[ch, q * 4 + i]]
                if cls:
                    tot += numpy.bincount(cls, minlength=8).max()
                    acc += 1
    return tot, acc


if __name__ == "__main__":
    c = slot_counts(int(sys.argv[1]) if len(sys.argv) > 1 else 22)[: 32 * 600]
    unit = numpy.concatenate([[0], numpy.cumsum(c)[:-1]]) // 2  # 16-byte unit of every slot start
    a, b = one_lane_per_slot(unit, c // 2), two_lanes_per_slot(unit, c // 2)
    print(f"one lane per slot : {a[0] / a[1]:.2f} wavefronts per quarter-warp load, {a[0]} in total")
    print(f"two lanes per slot: {b[0] / b[1]:.2f} wavefronts per quarter-warp load, {b[0]} in total")
