// host_tables.h -- host-side construction of the constant tables the kernels read
#pragma once
#include <stdint.h>
#include <string.h>

#include "a26_compiled.cuh"

namespace ngp_host {

// Stella TIA::dumpedInputPort / Paddles::update [3P-recall]: resistance = 1.4e6 * charge/4096,
// CPU cycles until INPTx bit 7 rises = 1.216e-6 * resistance * scanlines(262) * framerate(59.92)
inline void build_paddle_table(uint32_t *t)
{
    for (int c = 0; c <= a26::TRIGMAX; ++c) {
        int32_t resistance = (int32_t)(1400000 * (c / (float)a26::TRIGMAX));
        t[c] = (uint32_t)(1.216e-6 * resistance * 262.0 * 59.92f);
    }
}

// rom words + decode table + per-palette-entry colour-match weights: 2 bits per target colour =
// number of RGB channels of the palette entry equal to the target's (the reference matches per
// channel, utils.py:62)
inline void build_tables(a26::Tables &tables, const uint8_t *rom, const uint8_t ball[3], const uint8_t left[3],
                         const uint8_t right[3], const uint32_t palette[128])
{
    memset(&tables, 0, sizeof(tables));
    memcpy(tables.rom, rom, 2048);
    a26::DecodeTable dt;
    a26::build_decode_table(dt);
    memcpy(tables.decode, dt.e, sizeof(dt.e));
    a26::fill_blockmap(tables.blockmap);
    const uint8_t *targets[3] = {ball, left, right};
    for (int i = 0; i < 128; ++i) {
        uint32_t c = palette[i];
        uint8_t rgb[3] = {(uint8_t)(c >> 16), (uint8_t)(c >> 8), (uint8_t)c};
        uint8_t w = 0;
        for (int t = 0; t < 3; ++t) {
            int m = 0;
            for (int ch = 0; ch < 3; ++ch) m += rgb[ch] == targets[t][ch];
            w |= (uint8_t)(m << (2 * t));
        }
        tables.weight[i] = w;
    }
}

}  // namespace ngp_host
