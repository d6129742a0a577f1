// hive_tables.h -- host-side builder of the constant geometry tables the environment kernels read through
// EnvArgs::hop_lines (one device allocation; also built by the CPU emulator under tests/emu):
//   [GEO_HOP , +720)  144 x 5 words: is_straight_line masks of every cell (move_checker.py:249-265, raw deltas)
//   [GEO_RANK, +144)  per cell: rank of each of its six neighbours in tile.adjacent_tiles order, 3 bits per direction
//                     (board_tiles order: q descending, then r ascending; tile.py:111-123, :156-203)
//   [GEO_NBR , +288)  per cell two words: the six neighbour cells as bytes (d0..d3 | d4, d5), torus of tile.py:114-121
//   [GEO_NBRMASK, +720) 144 x 5 words: the six neighbours of every cell as a board
// (offsets GEO_* are defined next to the device code that uses them, in hive_core.cuh; include that first)
#pragma once
#include <stdint.h>
#include <vector>

namespace hive {

// direction ring d0=(+1,0) d1=(+1,+1) d2=(0,+1) d3=(-1,0) d4=(-1,-1) d5=(0,-1) on the 12 x 12 torus, cell = q*12+r
inline int table_nbr(int c, int i) {
    static const int dq[6] = {1, 1, 0, -1, -1, 0}, dr[6] = {0, 1, 1, 0, -1, -1};
    const int q = (c / 12 + dq[i] + 12) % 12, r = (c % 12 + dr[i] + 12) % 12;
    return q * 12 + r;
}

inline void build_geometry_tables(std::vector<uint32_t>& t) {
    t.assign(GEO_WORDS, 0);
    for (int o = 0; o < 144; o++)
        for (int x = 0; x < 144; x++) {
            int q1 = o / 12, r1 = o % 12, q2 = x / 12, r2 = x % 12;
            int d1 = q1 - q2, d2 = 12 - d1, dx = d1 < d2 ? d1 : d2;
            d1 = r1 - r2; d2 = 12 - d1;
            int dy = d1 < d2 ? d1 : d2;
            if (q1 == q2 || r1 == r2 || dy == dx) t[GEO_HOP + o * 5 + (x >> 5)] |= 1u << (x & 31);
        }
    for (int c = 0; c < 144; c++) {
        int nb[6], key[6];
        for (int i = 0; i < 6; i++) { nb[i] = table_nbr(c, i); key[i] = (11 - nb[i] / 12) * 12 + nb[i] % 12; }
        uint32_t ranks = 0;
        for (int i = 0; i < 6; i++) {
            int rank = 0;
            for (int j = 0; j < 6; j++) rank += key[j] < key[i];
            ranks |= (uint32_t)rank << (3 * i);
        }
        t[GEO_RANK + c] = ranks;
        t[GEO_NBR + 2 * c] = (uint32_t)nb[0] | ((uint32_t)nb[1] << 8) | ((uint32_t)nb[2] << 16) | ((uint32_t)nb[3] << 24);
        t[GEO_NBR + 2 * c + 1] = (uint32_t)nb[4] | ((uint32_t)nb[5] << 8);
        for (int i = 0; i < 6; i++) t[GEO_NBRMASK + c * 5 + (nb[i] >> 5)] |= 1u << (nb[i] & 31);
    }
}

}  // namespace hive
