#!/usr/bin/env python3
"""Shared-memory / traffic / recompute budget of a 2.5D z-marching tile kernel that fuses T Jacobi sweeps per
launch.  A MODEL, not a measurement: it makes the capacity argument checkable.

Per (x, y) tile of BX x BY output voxels marching along z, stage j = 1..T updates plane z0 - (j-1) on the tile
grown by (T - j) voxels per side; per plane position the kernel must hold in shared memory

  centre-only statics  fx, fy, fz, ft, ksi      5 fields x T planes            (one plane per stage)
  stencil statics      u, v, w, phi             4 fields x (T + 2) planes      (z-1 .. z+1 of every stage)
  iterates             du, dv, dw               3 fields x 3 planes x T        (input of every stage; the last
                                                                                stage's output goes to global memory)
  + P planes in flight for each of the 12 fields that come from global memory (asynchronous pipeline depth)

every buffer sized for the haloed box (BX + 2T) x (BY + 2T) floats (statics and the first stage's input) or the
stage's own box.  Global traffic per output voxel and sweep: 12 fields read on the haloed box + 3 written on the
tile, divided by T.  Recompute: sum of the stage boxes over T tiles.
"""
import itertools

SMEM_LIMIT = 227 * 1024


def budget(bx, by, T, P):
    hx, hy = bx + 2 * T, by + 2 * T
    box = hx * hy * 4
    planes = 5 * T + 4 * (T + 2) + 3 * 3 * T + 12 * P
    smem = planes * box
    words = (12.0 * hx * hy + 3.0 * bx * by) / (T * bx * by)
    recompute = sum((bx + 2 * (T - j)) * (by + 2 * (T - j)) for j in range(1, T + 1)) / float(T * bx * by)
    return planes, smem, words, recompute


def main():
    print("# T  P  tile     plane-buffers  smem KiB  CTAs/SM  words/voxel-sweep  recompute  (15 words = today's kernel)")
    for T, P in itertools.product((1, 2, 3, 5), (1, 3)):
        best = None
        for bx, by in itertools.product((16, 32, 64, 128), (4, 8, 16, 32)):
            planes, smem, words, rec = budget(bx, by, T, P)
            if smem > SMEM_LIMIT:
                continue
            # cost proxy: the kernel is balanced between DRAM traffic and instruction issue, so a
            # tile is as good as max(traffic relative to 15 words, instructions relative to one plain sweep)
            score = max(words / 15.0, rec)
            if best is None or score < best[0]:
                best = (score, bx, by, planes, smem, words, rec)
        if best is None:
            print("%d  %d  nothing fits in 227 KiB" % (T, P))
            continue
        _, bx, by, planes, smem, words, rec = best
        print("%d  %d  %3dx%-3d  %4d           %6.1f    %d        %5.2f              %.2fx" %
              (T, P, bx, by, planes, smem / 1024.0, SMEM_LIMIT // smem, words, rec))


if __name__ == "__main__":
    main()
