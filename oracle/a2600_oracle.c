/* a2600_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).  See a2600_oracle.h.
 *
 * Style: clarity first.  The TIA is stepped one colour clock at a time, the CPU is a plain
 * opcode switch.  The CUDA core (csrc/) uses a different algorithm (160-bit span masks per
 * register-constant segment, table-driven decode); bit-exact agreement between the two on
 * RAM, registers, collision latches and every framebuffer pixel is the parity test.
 *
 * Replaces, for testing only: retro.make/env.reset/env.step  (/root/reference/main.py:21,40,51,56,77,108).
 */
#include "a2600_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ---- Stella-convention constants (each is a named, swappable choice) ------- */
#define HBLANK_CLOCKS 68
#define LINE_CLOCKS 228
#define LINE_CYCLES 76
#define PADDLE_MAX_RESISTANCE 1400000
#define PADDLE_DIGITAL_SENSITIVITY 5   /* frames of 1,2,..5 steps before the fast step   */
#define PADDLE_DIGITAL_DISTANCE 60     /* fast step, charge units per frame               */
#define PADDLE_SCANLINES 262.0
#define PADDLE_FRAMERATE 59.92f
#define FRAME_CYCLE_CAP (4 * 262 * 76) /* safety stop if the program never ends VSYNC    */

enum { ERR_ILLEGAL_OPCODE = -1, ERR_DECIMAL = -2, ERR_RSYNC = -3 };

struct a26o {
    /* CPU */
    uint8_t a, x, y, sp, p;
    uint16_t pc;
    uint8_t dbus;
    uint8_t ram[128];
    uint8_t rom[2048];
    int64_t cycles;          /* CPU cycles since power-on; line grid is cycles % 76 */
    uint64_t instructions;
    int error;
    /* RIOT */
    int32_t timer_value;     /* value << shift at the time of the write */
    int32_t timer_shift;
    int64_t timer_set_cycle;
    uint8_t swcha, swchb;
    /* paddles */
    int32_t charge[4], repeat[4], keyrep[4];
    uint32_t needed_tab[A26O_TRIGMAX + 1];
    int dump_enabled;
    int64_t dump_disabled_cycle;
    /* TIA registers */
    uint8_t vsync, vblank;
    uint8_t nusiz0, nusiz1, colup0, colup1, colupf, colubk, ctrlpf, refp0, refp1;
    uint8_t pf0, pf1, pf2;
    uint8_t grp0_new, grp0_old, grp1_new, grp1_old;
    uint8_t enam0, enam1, enabl_new, enabl_old;
    uint8_t hmp0, hmp1, hmm0, hmm1, hmbl;
    uint8_t vdelp0, vdelp1, vdelbl, resmp0, resmp1;
    int16_t posp0, posp1, posm0, posm1, posbl;   /* 0..159: pixel where the object starts */
    uint8_t suppress_p0, suppress_p1;            /* RESPx hit this line: main copy not drawn */
    uint8_t hmove_blank;                         /* this line has the 8-pixel HMOVE comb */
    uint16_t cx;                                 /* 15 collision latches (bit layout below) */
    /* rendering */
    int64_t tia_clock;           /* next colour clock to be rendered (absolute) */
    int64_t frame_start_line;    /* absolute line index where the current frame started */
    int frame_done;
    uint8_t *fb;
};

/* collision latch bits: index = 2*reg + (D7 ? 1 : 0) for reg = CXM0P..CXPPMM (0..7) */
enum { CX_M0P0 = 0, CX_M0P1 = 1, CX_M1P1 = 2, CX_M1P0 = 3, CX_P0BL = 4, CX_P0PF = 5,
       CX_P1BL = 6, CX_P1PF = 7, CX_M0BL = 8, CX_M0PF = 9, CX_M1BL = 10, CX_M1PF = 11,
       CX_BLPF = 13, CX_M0M1 = 14, CX_P0P1 = 15 };

const uint32_t a26o_ntsc_palette[128] = {
    0x000000, 0x4a4a4a, 0x6f6f6f, 0x8e8e8e, 0xaaaaaa, 0xc0c0c0, 0xd6d6d6, 0xececec,
    0x484800, 0x69690f, 0x86861d, 0xa2a22a, 0xbbbb35, 0xd2d240, 0xe8e84a, 0xfcfc54,
    0x7c2c00, 0x904811, 0xa26221, 0xb47a30, 0xc3903d, 0xd2a44a, 0xdfb755, 0xecc860,
    0x901c00, 0xa33915, 0xb55328, 0xc66c3a, 0xd5824a, 0xe39759, 0xf0aa67, 0xfcbc74,
    0x940000, 0xa71a1a, 0xb83232, 0xc84848, 0xd65c5c, 0xe46f6f, 0xf08080, 0xfc9090,
    0x840064, 0x97197a, 0xa8308f, 0xb846a2, 0xc659b3, 0xd46cc3, 0xe07cd2, 0xec8ce0,
    0x500084, 0x68199a, 0x7d30ad, 0x9246c0, 0xa459d0, 0xb56ce0, 0xc57cee, 0xd48cfc,
    0x140090, 0x331aa3, 0x4e32b5, 0x6848c6, 0x7f5cd5, 0x956fe3, 0xa980f0, 0xbc90fc,
    0x000094, 0x181aa7, 0x2d32b8, 0x4248c8, 0x545cd6, 0x656fe4, 0x7580f0, 0x8490fc,
    0x001c88, 0x183b9d, 0x2d57b0, 0x4272c2, 0x548ad2, 0x65a0e1, 0x75b5ef, 0x84c8fc,
    0x003064, 0x185080, 0x2d6d98, 0x4288b0, 0x54a0c5, 0x65b7d9, 0x75cceb, 0x84e0fc,
    0x004030, 0x18624e, 0x2d8169, 0x429e82, 0x54b899, 0x65d1ae, 0x75e7c2, 0x84fcd4,
    0x004400, 0x1a661a, 0x328432, 0x48a048, 0x5cba5c, 0x6fd26f, 0x80e880, 0x90fc90,
    0x143c00, 0x355f18, 0x527e2d, 0x6e9c42, 0x87b754, 0x9ed065, 0xb4e775, 0xc8fc84,
    0x303800, 0x505916, 0x6d762b, 0x88923e, 0xa0ab4f, 0xb7c25f, 0xccd86e, 0xe0ec7c,
    0x482c00, 0x694d14, 0x866a26, 0xa28638, 0xbb9f47, 0xd2b656, 0xe8cc63, 0xfce070,
};

void a26o_fb_to_rgb(const uint8_t *fb, uint8_t *rgb)
{
    for (int i = 0; i < A26O_FB_ROWS * A26O_FB_COLS; ++i) {
        uint32_t c = a26o_ntsc_palette[fb[i] >> 1];
        rgb[3 * i + 0] = (uint8_t)(c >> 16);
        rgb[3 * i + 1] = (uint8_t)(c >> 8);
        rgb[3 * i + 2] = (uint8_t)c;
    }
}

void a26o_build_paddle_table(uint32_t needed[A26O_TRIGMAX + 1])
{
    for (int c = 0; c <= A26O_TRIGMAX; ++c) {
        int32_t resistance = (int32_t)(PADDLE_MAX_RESISTANCE * (c / (float)A26O_TRIGMAX));
        needed[c] = (uint32_t)(1.216e-6 * resistance * PADDLE_SCANLINES * PADDLE_FRAMERATE);
    }
}

/* ======================================= TIA ========================================== */

static int pf_bit(const a26o *s, int x)
{
    int idx;
    if (x < 80) idx = x >> 2;
    else idx = (s->ctrlpf & 1) ? (159 - x) >> 2 : (x - 80) >> 2;
    if (idx < 4) return (s->pf0 >> (4 + idx)) & 1;
    if (idx < 12) return (s->pf1 >> (11 - idx)) & 1;
    return (s->pf2 >> (idx - 12)) & 1;
}

static const int8_t copy_offsets[8][3] = {
    {0, -1, -1}, {0, 16, -1}, {0, 32, -1}, {0, 16, 32}, {0, 64, -1}, {0, -1, -1}, {0, 32, 64}, {0, -1, -1}};

static int player_bit(int x, int pos, int nusiz, uint8_t grp, int reflect, int suppress)
{
    int mode = nusiz & 7;
    int scale = mode == 5 ? 2 : mode == 7 ? 4 : 1;
    int d = (x - pos + 160) % 160;
    if (grp == 0) return 0;
    for (int c = 0; c < 3; ++c) {
        int off = copy_offsets[mode][c];
        if (off < 0) break;
        if (c == 0 && suppress) continue;
        int dd = d - off - (scale > 1 ? 1 : 0);   /* stretched players start one pixel late */
        if (dd >= 0 && dd < 8 * scale) {
            int bit = dd / scale;
            return (grp >> (reflect ? bit : 7 - bit)) & 1;
        }
    }
    return 0;
}

static int missile_bit(int x, int pos, int nusiz)
{
    int mode = nusiz & 7;
    int width = 1 << ((nusiz >> 4) & 3);
    int d = (x - pos + 160) % 160;
    for (int c = 0; c < 3; ++c) {
        int off = copy_offsets[mode][c];
        if (off < 0) break;
        if (d - off >= 0 && d - off < width) return 1;
    }
    return 0;
}

static void render_clock(a26o *s, int64_t clk)
{
    int h = (int)(clk % LINE_CLOCKS);
    int64_t line = clk / LINE_CLOCKS - s->frame_start_line;
    if (h == 0) {           /* new scanline: per-line latches drop */
        s->hmove_blank = 0;
        s->suppress_p0 = s->suppress_p1 = 0;
    }
    if (h < HBLANK_CLOCKS) return;
    if (line < A26O_YSTART || line >= A26O_YSTART + A26O_FB_ROWS) return;  /* outside display window */
    int x = h - HBLANK_CLOCKS;
    uint8_t colour = 0;
    if (!(s->vblank & 2)) {
        int pf = pf_bit(s, x);
        int bl = ((s->vdelbl & 1) ? s->enabl_old : s->enabl_new) & 2
                     ? ((x - s->posbl + 160) % 160) < (1 << ((s->ctrlpf >> 4) & 3)) : 0;
        int p0 = player_bit(x, s->posp0, s->nusiz0, (s->vdelp0 & 1) ? s->grp0_old : s->grp0_new,
                            s->refp0 & 8, s->suppress_p0);
        int p1 = player_bit(x, s->posp1, s->nusiz1, (s->vdelp1 & 1) ? s->grp1_old : s->grp1_new,
                            s->refp1 & 8, s->suppress_p1);
        int m0 = ((s->enam0 & 2) && !(s->resmp0 & 2)) ? missile_bit(x, s->posm0, s->nusiz0) : 0;
        int m1 = ((s->enam1 & 2) && !(s->resmp1 & 2)) ? missile_bit(x, s->posm1, s->nusiz1) : 0;
        uint16_t cx = 0;
        if (m0 && p0) cx |= 1u << CX_M0P0;
        if (m0 && p1) cx |= 1u << CX_M0P1;
        if (m1 && p1) cx |= 1u << CX_M1P1;
        if (m1 && p0) cx |= 1u << CX_M1P0;
        if (p0 && bl) cx |= 1u << CX_P0BL;
        if (p0 && pf) cx |= 1u << CX_P0PF;
        if (p1 && bl) cx |= 1u << CX_P1BL;
        if (p1 && pf) cx |= 1u << CX_P1PF;
        if (m0 && bl) cx |= 1u << CX_M0BL;
        if (m0 && pf) cx |= 1u << CX_M0PF;
        if (m1 && bl) cx |= 1u << CX_M1BL;
        if (m1 && pf) cx |= 1u << CX_M1PF;
        if (bl && pf) cx |= 1u << CX_BLPF;
        if (m0 && m1) cx |= 1u << CX_M0M1;
        if (p0 && p1) cx |= 1u << CX_P0P1;
        s->cx |= cx;
        /* priority encoder (Stella 3.x: with PF priority the score bit is not used) */
        if (s->ctrlpf & 4) {
            if (pf || bl) colour = s->colupf;
            else if (p0 || m0) colour = s->colup0;
            else if (p1 || m1) colour = s->colup1;
            else colour = s->colubk;
        } else {
            if (p0 || m0) colour = s->colup0;
            else if (p1 || m1) colour = s->colup1;
            else if (bl) colour = s->colupf;
            else if (pf) colour = (s->ctrlpf & 2) ? (x < 80 ? s->colup0 : s->colup1) : s->colupf;
            else colour = s->colubk;
        }
        if (s->hmove_blank && x < 8) colour = 0;
    }
    if (s->fb) s->fb[(line - A26O_YSTART) * A26O_FB_COLS + x] = colour & 0xFE;
}

static void tia_update(a26o *s, int64_t target)
{
    while (s->tia_clock < target) render_clock(s, s->tia_clock++);
}

static int hm_signed(uint8_t hm) { int v = hm >> 4; return v >= 8 ? v - 16 : v; }
static int wrap160(int v) { v %= 160; return v < 0 ? v + 160 : v; }

static void tia_poke(a26o *s, int reg, uint8_t v)
{
    int64_t clock = s->cycles * 3;                 /* the write cycle has already elapsed */
    int hpos = (int)(clock % LINE_CLOCKS);
    int delay = 0;
    switch (reg) {
    case 0x01: delay = 1; break;                                   /* VBLANK */
    case 0x04: case 0x05: delay = 8; break;                        /* NUSIZx */
    case 0x0B: case 0x0C: delay = 1; break;                        /* REFPx */
    case 0x0D: case 0x0E: case 0x0F: {                             /* PFx */
        static const int d[4] = {4, 5, 2, 3};
        delay = d[(hpos / 3) & 3];
        break;
    }
    case 0x1B: case 0x1C: case 0x1D: case 0x1E: case 0x1F: delay = 1; break;  /* GRPx ENAxx */
    default: break;
    }
    tia_update(s, clock + delay);
    switch (reg) {
    case 0x00:                                                     /* VSYNC */
        if ((s->vsync & 2) && !(v & 2)) {
            s->frame_done = 1;
            s->frame_start_line = clock / LINE_CLOCKS;
        }
        s->vsync = v;
        break;
    case 0x01:                                                     /* VBLANK */
        if (!(s->vblank & 0x80) && (v & 0x80)) s->dump_enabled = 1;
        if ((s->vblank & 0x80) && !(v & 0x80)) { s->dump_enabled = 0; s->dump_disabled_cycle = s->cycles; }
        s->vblank = v;
        break;
    case 0x02: {                                                   /* WSYNC */
        int c = (int)(s->cycles % LINE_CYCLES);
        if (c) s->cycles += LINE_CYCLES - c;
        break;
    }
    case 0x03: break;                    /* RSYNC: ignored (Stella 3.x does not emulate it) */
    case 0x04: s->nusiz0 = v; break;
    case 0x05: s->nusiz1 = v; break;
    case 0x06: s->colup0 = v; break;
    case 0x07: s->colup1 = v; break;
    case 0x08: s->colupf = v; break;
    case 0x09: s->colubk = v; break;
    case 0x0A: s->ctrlpf = v; break;
    case 0x0B: s->refp0 = v; break;
    case 0x0C: s->refp1 = v; break;
    case 0x0D: s->pf0 = v; break;
    case 0x0E: s->pf1 = v; break;
    case 0x0F: s->pf2 = v; break;
    case 0x10: s->posp0 = hpos < HBLANK_CLOCKS ? 3 : (hpos - HBLANK_CLOCKS + 5) % 160; s->suppress_p0 = 1; break;
    case 0x11: s->posp1 = hpos < HBLANK_CLOCKS ? 3 : (hpos - HBLANK_CLOCKS + 5) % 160; s->suppress_p1 = 1; break;
    case 0x12: s->posm0 = hpos < HBLANK_CLOCKS ? 2 : (hpos - HBLANK_CLOCKS + 4) % 160; break;
    case 0x13: s->posm1 = hpos < HBLANK_CLOCKS ? 2 : (hpos - HBLANK_CLOCKS + 4) % 160; break;
    case 0x14: s->posbl = hpos < HBLANK_CLOCKS ? 2 : (hpos - HBLANK_CLOCKS + 4) % 160; break;
    case 0x1B: s->grp0_new = v; s->grp1_old = s->grp1_new; break;
    case 0x1C: s->grp1_new = v; s->grp0_old = s->grp0_new; s->enabl_old = s->enabl_new; break;
    case 0x1D: s->enam0 = v; break;
    case 0x1E: s->enam1 = v; break;
    case 0x1F: s->enabl_new = v; break;
    case 0x20: s->hmp0 = v; break;
    case 0x21: s->hmp1 = v; break;
    case 0x22: s->hmm0 = v; break;
    case 0x23: s->hmm1 = v; break;
    case 0x24: s->hmbl = v; break;
    case 0x25: s->vdelp0 = v; break;
    case 0x26: s->vdelp1 = v; break;
    case 0x27: s->vdelbl = v; break;
    case 0x28:
        if ((s->resmp0 & 2) && !(v & 2)) {
            int mode = s->nusiz0 & 7;
            s->posm0 = wrap160(s->posp0 + (mode == 5 ? 8 : mode == 7 ? 16 : 4));
        }
        s->resmp0 = v;
        break;
    case 0x29:
        if ((s->resmp1 & 2) && !(v & 2)) {
            int mode = s->nusiz1 & 7;
            s->posm1 = wrap160(s->posp1 + (mode == 5 ? 8 : mode == 7 ? 16 : 4));
        }
        s->resmp1 = v;
        break;
    case 0x2A:                                                     /* HMOVE */
        /* Standard model: every object moves by its signed HM nibble (positive = left);
         * a strobe in the first 21 cycles of the line blanks the first 8 pixels. */
        s->posp0 = wrap160(s->posp0 - hm_signed(s->hmp0));
        s->posp1 = wrap160(s->posp1 - hm_signed(s->hmp1));
        s->posm0 = wrap160(s->posm0 - hm_signed(s->hmm0));
        s->posm1 = wrap160(s->posm1 - hm_signed(s->hmm1));
        s->posbl = wrap160(s->posbl - hm_signed(s->hmbl));
        if (hpos / 3 <= 20) s->hmove_blank = 1;
        break;
    case 0x2B: s->hmp0 = s->hmp1 = s->hmm0 = s->hmm1 = s->hmbl = 0; break;
    case 0x2C: s->cx = 0; break;
    default: break;                                                /* audio & unused */
    }
}

static uint8_t inpt_read(a26o *s, int i)
{
    if (s->dump_enabled) return 0x00;
    uint32_t needed = s->needed_tab[s->charge[i]];
    return (uint64_t)(s->cycles - s->dump_disabled_cycle) > needed ? 0x80 : 0x00;
}

static uint8_t tia_peek(a26o *s, int reg)
{
    uint8_t noise = s->dbus & 0x3F;
    tia_update(s, s->cycles * 3);
    switch (reg & 0x0F) {
    case 0: case 1: case 2: case 3: case 4: case 5: case 6: case 7: {
        int r = reg & 7;
        uint8_t v = (uint8_t)((((s->cx >> (2 * r + 1)) & 1) << 7) | (((s->cx >> (2 * r)) & 1) << 6));
        if (r == 6) v &= 0x80;
        return v | noise;
    }
    case 8: case 9: case 10: case 11: return inpt_read(s, (reg & 0x0F) - 8) | noise;
    case 12: case 13: return 0x80 | noise;       /* joystick triggers: not pressed */
    default: return noise;
    }
}

/* ======================================= RIOT ========================================= */

static uint8_t riot_peek(a26o *s, uint16_t addr)
{
    switch (addr & 7) {
    case 0: return s->swcha;
    case 1: return 0x00;          /* SWACNT */
    case 2: return s->swchb;
    case 3: return 0x00;          /* SWBCNT */
    case 4: case 6: {             /* INTIM */
        int32_t t = s->timer_value - (int32_t)(s->cycles - s->timer_set_cycle);
        if (t >= 0) return (uint8_t)(t >> s->timer_shift);
        return (uint8_t)t;        /* after underflow: counts down once per cycle */
    }
    default: return 0x00;         /* TIMINT: unused by the cartridge */
    }
}

static void riot_poke(a26o *s, uint16_t addr, uint8_t v)
{
    if (addr & 4) {
        if (addr & 0x10) {
            static const int shift[4] = {0, 3, 6, 10};
            s->timer_shift = shift[addr & 3];
            s->timer_value = (int32_t)v << s->timer_shift;
            s->timer_set_cycle = s->cycles;
        }
    }
    /* SWCHA/SWACNT/SWCHB/SWBCNT writes: ignored (all pins are inputs here) */
}

/* ======================================= bus ========================================== */

static uint8_t rd(a26o *s, uint16_t addr)
{
    uint8_t v;
    addr &= 0x1FFF;
    if (addr & 0x1000) v = s->rom[addr & 0x7FF];
    else if (!(addr & 0x80)) v = tia_peek(s, addr & 0x0F);
    else if (!(addr & 0x200)) v = s->ram[addr & 0x7F];
    else v = riot_peek(s, addr);
    s->dbus = v;
    return v;
}

static void wr(a26o *s, uint16_t addr, uint8_t v)
{
    addr &= 0x1FFF;
    s->dbus = v;
    if (addr & 0x1000) return;
    if (!(addr & 0x80)) tia_poke(s, addr & 0x3F, v);
    else if (!(addr & 0x200)) s->ram[addr & 0x7F] = v;
    else riot_poke(s, addr, v);
}

/* ======================================= CPU ========================================== */

enum { FC = 1, FZ = 2, FI = 4, FD = 8, FB = 16, FU = 32, FV = 64, FN = 128 };

static void setnz(a26o *s, uint8_t v) { s->p = (uint8_t)((s->p & ~(FN | FZ)) | (v & 0x80) | (v ? 0 : FZ)); }
static uint8_t fetch(a26o *s) { return rd(s, s->pc++); }
static void push(a26o *s, uint8_t v) { wr(s, 0x100 | s->sp, v); s->sp--; }
static uint8_t pull(a26o *s) { s->sp++; return rd(s, 0x100 | s->sp); }

static void adc(a26o *s, uint8_t m)
{
    if (s->p & FD) s->error = ERR_DECIMAL;
    unsigned sum = s->a + m + (s->p & FC);
    s->p = (uint8_t)((s->p & ~(FC | FV)) | (sum > 0xFF ? FC : 0) | ((~(s->a ^ m) & (s->a ^ sum) & 0x80) ? FV : 0));
    s->a = (uint8_t)sum;
    setnz(s, s->a);
}
static void cmp(a26o *s, uint8_t r, uint8_t m)
{
    s->p = (uint8_t)((s->p & ~FC) | (r >= m ? FC : 0));
    setnz(s, (uint8_t)(r - m));
}

static void cpu_step(a26o *s)
{
    int64_t start = s->cycles;
    uint8_t op = fetch(s);
    int n = 2;                /* total cycles of this instruction */
    uint16_t ea = 0;
    int crossed = 0;
    s->instructions++;

/* addressing-mode helpers: compute ea; operand fetches have no timing side effects */
#define ZP()   do { ea = fetch(s); } while (0)
#define ZPX()  do { ea = (uint8_t)(fetch(s) + s->x); } while (0)
#define ZPY()  do { ea = (uint8_t)(fetch(s) + s->y); } while (0)
#define ABS()  do { uint8_t lo = fetch(s); ea = (uint16_t)(lo | (fetch(s) << 8)); } while (0)
#define ABX()  do { uint8_t lo = fetch(s); uint16_t b = (uint16_t)(lo | (fetch(s) << 8)); ea = (uint16_t)(b + s->x); crossed = (b ^ ea) >> 8 != 0; } while (0)
#define ABY()  do { uint8_t lo = fetch(s); uint16_t b = (uint16_t)(lo | (fetch(s) << 8)); ea = (uint16_t)(b + s->y); crossed = (b ^ ea) >> 8 != 0; } while (0)
#define IZX()  do { uint8_t z = (uint8_t)(fetch(s) + s->x); uint8_t lo = rd(s, z); ea = (uint16_t)(lo | (rd(s, (uint8_t)(z + 1)) << 8)); } while (0)
#define IZY()  do { uint8_t z = fetch(s); uint8_t lo = rd(s, z); uint16_t b = (uint16_t)(lo | (rd(s, (uint8_t)(z + 1)) << 8)); ea = (uint16_t)(b + s->y); crossed = (b ^ ea) >> 8 != 0; } while (0)
/* the data access happens on the last cycle: devices see the cycle count after it */
#define LOAD(nc) (n = (nc) + crossed, s->cycles = start + n, rd(s, ea))
#define STORE(nc, v) do { n = (nc); s->cycles = start + n; wr(s, ea, (v)); } while (0)
#define RMW(nc, expr) do { n = (nc); s->cycles = start + n - 2; uint8_t m = rd(s, ea); uint8_t r; expr; s->cycles = start + n; wr(s, ea, r); setnz(s, r); } while (0)
#define BRANCH(cond) do { int8_t off = (int8_t)fetch(s); if (cond) { uint16_t t = (uint16_t)(s->pc + off); n = 3 + (((t ^ s->pc) & 0xFF00) ? 1 : 0); s->pc = t; } } while (0)
#define ASL_(m) (s->p = (uint8_t)((s->p & ~FC) | ((m) >> 7)), r = (uint8_t)((m) << 1))
#define LSR_(m) (s->p = (uint8_t)((s->p & ~FC) | ((m) & 1)), r = (uint8_t)((m) >> 1))
#define ROL_(m) do { uint8_t c = s->p & FC; s->p = (uint8_t)((s->p & ~FC) | ((m) >> 7)); r = (uint8_t)(((m) << 1) | c); } while (0)
#define ROR_(m) do { uint8_t c = s->p & FC; s->p = (uint8_t)((s->p & ~FC) | ((m) & 1)); r = (uint8_t)(((m) >> 1) | (c << 7)); } while (0)
#define ALU8(base, OPER) \
    case base + 0x01: IZX(); { uint8_t m = LOAD(6); OPER; } break; \
    case base + 0x05: ZP();  { uint8_t m = LOAD(3); OPER; } break; \
    case base + 0x09: { uint8_t m = fetch(s); n = 2; OPER; } break; \
    case base + 0x0D: ABS(); { uint8_t m = LOAD(4); OPER; } break; \
    case base + 0x11: IZY(); { uint8_t m = LOAD(5); OPER; } break; \
    case base + 0x15: ZPX(); { uint8_t m = LOAD(4); OPER; } break; \
    case base + 0x19: ABY(); { uint8_t m = LOAD(4); OPER; } break; \
    case base + 0x1D: ABX(); { uint8_t m = LOAD(4); OPER; } break;
#define SHIFT5(base, OPER) \
    case base + 0x06: ZP();  RMW(5, OPER); break; \
    case base + 0x0A: { uint8_t m = s->a; uint8_t r; OPER; s->a = r; setnz(s, r); n = 2; } break; \
    case base + 0x0E: ABS(); RMW(6, OPER); break; \
    case base + 0x16: ZPX(); RMW(6, OPER); break; \
    case base + 0x1E: ABX(); crossed = 0; RMW(7, OPER); break;

    switch (op) {
    ALU8(0x00, (s->a |= m, setnz(s, s->a)))
    ALU8(0x20, (s->a &= m, setnz(s, s->a)))
    ALU8(0x40, (s->a ^= m, setnz(s, s->a)))
    ALU8(0x60, adc(s, m))
    ALU8(0xA0, (s->a = m, setnz(s, s->a)))
    ALU8(0xC0, cmp(s, s->a, m))
    ALU8(0xE0, adc(s, (uint8_t)~m))
    SHIFT5(0x00, ASL_(m))
    SHIFT5(0x20, ROL_(m))
    SHIFT5(0x40, LSR_(m))
    SHIFT5(0x60, ROR_(m))
    /* STA */
    case 0x81: IZX(); STORE(6, s->a); break;
    case 0x85: ZP();  STORE(3, s->a); break;
    case 0x8D: ABS(); STORE(4, s->a); break;
    case 0x91: IZY(); STORE(6, s->a); break;
    case 0x95: ZPX(); STORE(4, s->a); break;
    case 0x99: ABY(); STORE(5, s->a); break;
    case 0x9D: ABX(); STORE(5, s->a); break;
    /* STX / STY */
    case 0x86: ZP();  STORE(3, s->x); break;
    case 0x96: ZPY(); STORE(4, s->x); break;
    case 0x8E: ABS(); STORE(4, s->x); break;
    case 0x84: ZP();  STORE(3, s->y); break;
    case 0x94: ZPX(); STORE(4, s->y); break;
    case 0x8C: ABS(); STORE(4, s->y); break;
    /* LDX / LDY */
    case 0xA2: s->x = fetch(s); setnz(s, s->x); break;
    case 0xA6: ZP();  s->x = LOAD(3); setnz(s, s->x); break;
    case 0xB6: ZPY(); s->x = LOAD(4); setnz(s, s->x); break;
    case 0xAE: ABS(); s->x = LOAD(4); setnz(s, s->x); break;
    case 0xBE: ABY(); s->x = LOAD(4); setnz(s, s->x); break;
    case 0xA0: s->y = fetch(s); setnz(s, s->y); break;
    case 0xA4: ZP();  s->y = LOAD(3); setnz(s, s->y); break;
    case 0xB4: ZPX(); s->y = LOAD(4); setnz(s, s->y); break;
    case 0xAC: ABS(); s->y = LOAD(4); setnz(s, s->y); break;
    case 0xBC: ABX(); s->y = LOAD(4); setnz(s, s->y); break;
    /* CPX / CPY */
    case 0xE0: cmp(s, s->x, fetch(s)); break;
    case 0xE4: ZP();  cmp(s, s->x, LOAD(3)); break;
    case 0xEC: ABS(); cmp(s, s->x, LOAD(4)); break;
    case 0xC0: cmp(s, s->y, fetch(s)); break;
    case 0xC4: ZP();  cmp(s, s->y, LOAD(3)); break;
    case 0xCC: ABS(); cmp(s, s->y, LOAD(4)); break;
    /* BIT */
    case 0x24: ZP();  { uint8_t m = LOAD(3); s->p = (uint8_t)((s->p & ~(FN | FV | FZ)) | (m & 0xC0) | ((m & s->a) ? 0 : FZ)); } break;
    case 0x2C: ABS(); { uint8_t m = LOAD(4); s->p = (uint8_t)((s->p & ~(FN | FV | FZ)) | (m & 0xC0) | ((m & s->a) ? 0 : FZ)); } break;
    /* INC / DEC memory */
    case 0xE6: ZP();  RMW(5, r = (uint8_t)(m + 1)); break;
    case 0xF6: ZPX(); RMW(6, r = (uint8_t)(m + 1)); break;
    case 0xEE: ABS(); RMW(6, r = (uint8_t)(m + 1)); break;
    case 0xFE: ABX(); crossed = 0; RMW(7, r = (uint8_t)(m + 1)); break;
    case 0xC6: ZP();  RMW(5, r = (uint8_t)(m - 1)); break;
    case 0xD6: ZPX(); RMW(6, r = (uint8_t)(m - 1)); break;
    case 0xCE: ABS(); RMW(6, r = (uint8_t)(m - 1)); break;
    case 0xDE: ABX(); crossed = 0; RMW(7, r = (uint8_t)(m - 1)); break;
    /* branches */
    case 0x10: BRANCH(!(s->p & FN)); break;
    case 0x30: BRANCH(s->p & FN); break;
    case 0x50: BRANCH(!(s->p & FV)); break;
    case 0x70: BRANCH(s->p & FV); break;
    case 0x90: BRANCH(!(s->p & FC)); break;
    case 0xB0: BRANCH(s->p & FC); break;
    case 0xD0: BRANCH(!(s->p & FZ)); break;
    case 0xF0: BRANCH(s->p & FZ); break;
    /* jumps / subroutines */
    case 0x4C: ABS(); s->pc = ea; n = 3; break;
    case 0x6C: { ABS(); uint8_t lo = rd(s, ea); uint8_t hi = rd(s, (uint16_t)((ea & 0xFF00) | ((ea + 1) & 0xFF))); s->pc = (uint16_t)(lo | (hi << 8)); n = 5; } break;
    case 0x20: { uint8_t lo = fetch(s); s->cycles = start + 4; push(s, (uint8_t)(s->pc >> 8)); s->cycles = start + 5; push(s, (uint8_t)s->pc);
                 s->pc = (uint16_t)(lo | (fetch(s) << 8)); n = 6; } break;
    case 0x60: { uint8_t lo = pull(s); uint8_t hi = pull(s); s->pc = (uint16_t)((lo | (hi << 8)) + 1); n = 6; } break;
    case 0x00: { s->pc++; s->cycles = start + 3; push(s, (uint8_t)(s->pc >> 8)); s->cycles = start + 4; push(s, (uint8_t)s->pc);
                 s->cycles = start + 5; push(s, (uint8_t)(s->p | FB | FU)); s->p |= FI;
                 s->pc = (uint16_t)(s->rom[0x7FE] | (s->rom[0x7FF] << 8)); n = 7; } break;
    case 0x40: { s->p = (uint8_t)((pull(s) & ~FB) | FU); uint8_t lo = pull(s); uint8_t hi = pull(s); s->pc = (uint16_t)(lo | (hi << 8)); n = 6; } break;
    /* stack */
    case 0x48: n = 3; s->cycles = start + 3; push(s, s->a); break;
    case 0x08: n = 3; s->cycles = start + 3; push(s, (uint8_t)(s->p | FB | FU)); break;
    case 0x68: n = 4; s->cycles = start + 4; s->a = pull(s); setnz(s, s->a); break;
    case 0x28: n = 4; s->cycles = start + 4; s->p = (uint8_t)((pull(s) & ~FB) | FU); break;
    /* flags */
    case 0x18: s->p &= ~FC; break;
    case 0x38: s->p |= FC; break;
    case 0x58: s->p &= ~FI; break;
    case 0x78: s->p |= FI; break;
    case 0xB8: s->p &= ~FV; break;
    case 0xD8: s->p &= ~FD; break;
    case 0xF8: s->p |= FD; break;
    /* register ops */
    case 0xAA: s->x = s->a; setnz(s, s->x); break;
    case 0xA8: s->y = s->a; setnz(s, s->y); break;
    case 0x8A: s->a = s->x; setnz(s, s->a); break;
    case 0x98: s->a = s->y; setnz(s, s->a); break;
    case 0xBA: s->x = s->sp; setnz(s, s->x); break;
    case 0x9A: s->sp = s->x; break;
    case 0xE8: s->x++; setnz(s, s->x); break;
    case 0xC8: s->y++; setnz(s, s->y); break;
    case 0xCA: s->x--; setnz(s, s->x); break;
    case 0x88: s->y--; setnz(s, s->y); break;
    case 0xEA: break;
    default: s->error = ERR_ILLEGAL_OPCODE; break;
    }
    /* WSYNC may have pushed cycles past start+n already */
    if (s->cycles < start + n) s->cycles = start + n;
}

/* ======================================= API ========================================== */

a26o *a26o_new(const uint8_t rom[2048])
{
    a26o *s = (a26o *)calloc(1, sizeof(a26o));
    memcpy(s->rom, rom, 2048);
    a26o_build_paddle_table(s->needed_tab);
    a26o_power_on(s);
    return s;
}
void a26o_free(a26o *s) { free(s); }

void a26o_power_on(a26o *s)
{
    uint8_t rom[2048];
    uint32_t *tab = (uint32_t *)malloc(sizeof(s->needed_tab));
    memcpy(rom, s->rom, 2048);
    memcpy(tab, s->needed_tab, sizeof(s->needed_tab));
    memset(s, 0, sizeof(*s));
    memcpy(s->rom, rom, 2048);
    memcpy(s->needed_tab, tab, sizeof(s->needed_tab));
    free(tab);
    s->p = FU | FI;
    s->sp = 0xFD;
    s->pc = (uint16_t)(s->rom[0x7FC] | (s->rom[0x7FD] << 8));
    s->swcha = 0xFF;
    s->swchb = 0x3F;
    for (int i = 0; i < 4; ++i) s->charge[i] = A26O_TRIGMAX / 2;
    /* RIOT timer after power-on: arbitrary but fixed */
    s->timer_shift = 10;
    s->timer_value = 0;
    s->timer_set_cycle = 0;
    s->dump_enabled = 0;
    s->dump_disabled_cycle = 0;
}

static void paddles_update(a26o *s, const a26o_input *in)
{
    for (int i = 0; i < 4; ++i) {
        if (s->keyrep[i]) {
            s->repeat[i]++;
            if (s->repeat[i] > PADDLE_DIGITAL_SENSITIVITY) s->repeat[i] = PADDLE_DIGITAL_DISTANCE;
        }
        s->keyrep[i] = 0;
        if ((in->dec >> i) & 1) {
            s->keyrep[i] = 1;
            if (s->charge[i] > s->repeat[i]) s->charge[i] -= s->repeat[i];
        }
        if ((in->inc >> i) & 1) {
            s->keyrep[i] = 1;
            if (s->charge[i] + s->repeat[i] < A26O_TRIGMAX) s->charge[i] += s->repeat[i];
        }
    }
    /* SWCHA: paddle fire buttons, active low: P0=bit7, P1=bit6, P2=bit3, P3=bit2 */
    uint8_t a = 0xFF;
    if (in->fire & 1) a &= ~0x80;
    if (in->fire & 2) a &= ~0x40;
    if (in->fire & 4) a &= ~0x08;
    if (in->fire & 8) a &= ~0x04;
    s->swcha = a;
    s->swchb = in->swchb;
}

int a26o_run_frame(a26o *s, const a26o_input *in, uint8_t *fb)
{
    int64_t start = s->cycles;
    paddles_update(s, in);
    s->fb = fb;
    if (fb) memset(fb, 0, A26O_FB_ROWS * A26O_FB_COLS);
    s->frame_done = 0;
    while (!s->frame_done && !s->error && s->cycles - start < FRAME_CYCLE_CAP) cpu_step(s);
    /* bring the renderer up to the CPU's clock so the framebuffer is complete */
    tia_update(s, s->cycles * 3);
    s->fb = NULL;
    return s->error;
}

const uint8_t *a26o_ram(const a26o *s) { return s->ram; }
void a26o_cpu_regs(const a26o *s, uint8_t out[8])
{
    out[0] = s->a; out[1] = s->x; out[2] = s->y; out[3] = s->sp; out[4] = s->p;
    out[5] = (uint8_t)s->pc; out[6] = (uint8_t)(s->pc >> 8); out[7] = 0;
}
uint64_t a26o_cycles(const a26o *s) { return (uint64_t)s->cycles; }
uint64_t a26o_instructions(const a26o *s) { return s->instructions; }
int a26o_state_size(void) { return (int)sizeof(a26o); }
void a26o_save(const a26o *s, void *dst) { memcpy(dst, s, sizeof(*s)); }
void a26o_load(a26o *s, const void *src) { memcpy(s, src, sizeof(*s)); }

void a26o_tia_digest(const a26o *s, uint32_t out[8])
{
    out[0] = s->cx;
    out[1] = (uint32_t)(s->posp0 | (s->posp1 << 8) | (s->posm0 << 16) | (s->posm1 << 24));
    out[2] = (uint32_t)s->posbl | ((uint32_t)s->vblank << 8) | ((uint32_t)s->ctrlpf << 16) | ((uint32_t)s->vdelbl << 24);
    out[3] = (uint32_t)s->charge[0] | ((uint32_t)s->charge[1] << 16);
    out[4] = (uint32_t)s->charge[2] | ((uint32_t)s->charge[3] << 16);
    out[5] = (uint32_t)s->grp0_new | ((uint32_t)s->grp1_new << 8) | ((uint32_t)s->enabl_new << 16) | ((uint32_t)s->enabl_old << 24);
    out[6] = (uint32_t)(s->cycles % LINE_CYCLES);
    out[7] = (uint32_t)s->dump_enabled;
}
