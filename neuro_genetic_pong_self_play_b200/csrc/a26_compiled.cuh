// a26_compiled.cuh -- the 6507 side of the fused rollout as statically translated code.
//
// tools/gen_rom_core.py turns every reachable instruction of the bundled cartridge into specialised
// statements (generated/pong_core.inc).  This header provides the dispatcher around them.  TIA, RIOT,
// paddles, rendering and observation accumulation are the same device functions the interpreter
// (a26_core.cuh::run_frame, the verify-mode core) uses, so only instruction fetch/decode/addressing is
// specialised.  Control transfers return to `switch (blockmap[pc])`, where diverged lanes re-converge;
// the loop is scanline-synchronous like the interpreter's.  A program counter outside the translated set
// raises ERR_UNTRANSLATED (never happens for this cartridge: the traversal covers all of its code).
#pragma once
#include "a26_core.cuh"

namespace a26 {

enum : int { ERR_UNTRANSLATED = 4 };

#define A26_COMPILED_BLOCKMAP
#include "generated/pong_core.inc"
#undef A26_COMPILED_BLOCKMAP

inline void fill_blockmap(uint16_t *dst) { for (int i = 0; i < 2048; ++i) dst[i] = kCompiledBlockMap[i]; }

#ifdef __CUDACC__

#define A26_ADC(M_)                                              \
    do {                                                         \
        const uint32_t m2_ = (M_);                               \
        if (fid & 8u) { s.error = ERR_DECIMAL; done = 1; }       \
        const uint32_t sum_ = a + m2_ + fc;                      \
        fv = ((~(a ^ m2_) & (a ^ sum_)) >> 7) & 1u;              \
        fc = sum_ >> 8;                                          \
        a = sum_ & 0xFFu; nv = zv = a;                           \
    } while (0)
#define A26_CMP(R_, M_)                                          \
    do {                                                         \
        const uint32_t t_ = (R_) - (M_);                         \
        fc = (R_) >= (M_) ? 1u : 0u;                             \
        nv = zv = t_ & 0xFFu;                                    \
    } while (0)
#define A26_PACK_P() (fc | ((((zv & 0xFFu) == 0u) ? 1u : 0u) << 1) | fid | 0x20u | (fv << 6) | (nv & 0x80u))
#define A26_UNPACK_P(P_)                                         \
    do {                                                         \
        fc = (P_) & 1u; zv = ((P_) & 2u) ? 0u : 1u; fid = (P_) & 0x0Cu; fv = ((P_) >> 6) & 1u; nv = (P_) & 0x80u; \
    } while (0)
#define A26_WRITE_DYN(ADDR_, VAL_, TAFTER_)                                                           \
    do {                                                                                              \
        const uint32_t wa_ = (ADDR_) & 0x1FFFu;                                                       \
        if ((wa_ & 0x1280u) == 0x0080u) ram.wr(wa_, (VAL_));                                          \
        else if (!(wa_ & 0x1000u)) {                                                                  \
            const uint32_t rs_ = io_write_slow<VERIFY>(s, T, wa_, (VAL_), (TAFTER_), cpu_ls, fb);     \
            stall_ += rs_ & 0xFFFFu;                                                                  \
            if (rs_ >> 16) done = 1;                                                                  \
        }                                                                                             \
    } while (0)

template <bool VERIFY>
__device__ __forceinline__ void run_frame_compiled(Chip &s, CpuRegs &r, const Tables &T, Ram ram, uint8_t *fb)
{
    uint32_t a = r.a, x = r.x, y = r.y, sp = r.sp, pc = r.pc;
    uint32_t fc = r.c, fv = r.v, nv = r.nv, zv = r.zv, fid = r.id;
    uint32_t cyc = r.cyc, cpu_ls = r.cpu_ls;
    const uint32_t start_cyc = cyc;
    uint32_t done = 0;
    s.frame_done = 0;
    if (s.error) done = 1;
    while (!done && (cyc - start_cyc) < FRAME_CYCLE_CAP) {
        const uint32_t line_end = cpu_ls + LINE_CYCLES;
        while ((int32_t)(cyc - line_end) < 0 && !done) {
            A26_STAT(5);
            uint32_t entry;
            A26_HOT_DISPATCH
            entry = (pc & 0x1000u) ? T.blockmap[pc & 0x7FFu] : 0u;
            A26_STAT_ENTRY(pc);
            switch (entry) {
#include "generated/pong_core.inc"
            default:
                s.error = (uint8_t)((pc & 0x1000u) ? (int)ERR_UNTRANSLATED : (int)ERR_PC_NOT_ROM);
                done = 1;
                break;
            }
        a26_next_:;
        }
        while ((int32_t)(cyc - (cpu_ls + LINE_CYCLES)) >= 0) cpu_ls += LINE_CYCLES;
    }
    tia_catchup<VERIFY>(s, T, 3 * (int)(cyc - s.tia_ls), fb);
    r.a = a; r.x = x; r.y = y; r.sp = sp; r.pc = pc; r.c = fc; r.v = fv; r.nv = nv; r.zv = zv; r.id = fid;
    r.cyc = cyc; r.cpu_ls = cpu_ls;
}

#endif  // __CUDACC__
}  // namespace a26
