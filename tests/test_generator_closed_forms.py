"""The generator replaces two register-only loops of the cartridge by closed forms (tools/gen_rom_core.py:
closed_form_loop).  This test cuts the emitted statements out of csrc/generated/pong_core.inc, compiles them for the host
next to an instruction-by-instruction execution of the same loops (same A26_ADC macro, plain loop structure) and compares
registers, flags and cycle counts for every possible input."""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "neuro_genetic_pong_self_play_b200", "csrc", "generated", "pong_core.inc")
HDR = os.path.join(ROOT, "neuro_genetic_pong_self_play_b200", "csrc", "a26_compiled.cuh")


def _block(text, marker):
    i = text.index(marker)
    start = text.rindex("  { /*", 0, i)
    end = text.index("\n  }\n", i) + len("\n  }\n")
    return text[start:end]


def test_closed_form_loops_match_instruction_by_instruction_execution():
    inc = open(INC).read()
    dey = _block(inc, "/* DEY ; BPL F457: n iterations in closed form */")
    div = _block(inc, "/* INY ; SBC #$0F ; BCS F44A: closed form")
    adc = re.search(r"#define A26_ADC\(M_\).*?while \(0\)\n", open(HDR).read(), re.S).group(0)
    src = r'''
#include <stdint.h>
#include <stdio.h>
enum { ERR_DECIMAL = 2 };
struct S { int error; } s;
%s
struct R { uint32_t a, y, fc, fv, nv, zv, cyc; };
static R closed_dey(R r) { uint32_t a = r.a, y = r.y, fc = r.fc, fv = r.fv, nv = r.nv, zv = r.zv, cyc = r.cyc, fid = 4, done = 0; (void)fid; (void)done;
%s
  return R{a, y, fc, fv, nv, zv, cyc}; }
static R closed_div(R r) { uint32_t a = r.a, y = r.y, fc = r.fc, fv = r.fv, nv = r.nv, zv = r.zv, cyc = r.cyc, fid = 4, done = 0; (void)done;
%s
  return R{a, y, fc, fv, nv, zv, cyc}; }
static R loop_dey(R r) { uint32_t y = r.y, nv = r.nv, zv = r.zv, cyc = r.cyc;
  for (;;) { y = (y - 1) & 0xFFu; nv = zv = y; cyc += 2u; if (((nv >> 7) & 1u) == 0u) { cyc += 3u; continue; } cyc += 2u; break; }
  return R{r.a, y, r.fc, r.fv, nv, zv, cyc}; }
static R loop_div(R r) { uint32_t a = r.a, y = r.y, fc = r.fc, fv = r.fv, nv = r.nv, zv = r.zv, cyc = r.cyc, fid = 4, done = 0; (void)done;
  for (;;) { y = (y + 1) & 0xFFu; nv = zv = y; cyc += 2u; A26_ADC(0x0Fu ^ 0xFFu); cyc += 2u; if (fc == 1u) { cyc += 3u; continue; } cyc += 2u; break; }
  return R{a, y, fc, fv, nv, zv, cyc}; }
static int same(R p, R q) { return p.a == q.a && p.y == q.y && p.fc == q.fc && p.fv == q.fv && (p.nv & 0x80u) == (q.nv & 0x80u) &&
                                   ((p.zv & 0xFFu) == 0) == ((q.zv & 0xFFu) == 0) && p.cyc == q.cyc; }
int main() {
  int bad = 0;
  for (uint32_t y = 0; y < 256; ++y) { R r{0x5A, y, 1, 0, 0, 1, 1000}; if (!same(closed_dey(r), loop_dey(r))) { ++bad; printf("dey y=%%u\n", y); } }
  for (uint32_t a = 0; a < 256; ++a) for (uint32_t c = 0; c < 2; ++c) for (uint32_t y = 0; y < 256; y += 51) {
    R r{a, y, c, 0, 0, 1, 77}; if (!same(closed_div(r), loop_div(r))) { ++bad; printf("div a=%%u c=%%u y=%%u\n", a, c, y); } }
  printf("bad=%%d\n", bad);
  return bad != 0;
}
''' % (adc, dey, div)
    with tempfile.TemporaryDirectory() as d:
        cpp, exe = os.path.join(d, "t.cpp"), os.path.join(d, "t")
        open(cpp, "w").write(src)
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, cpp])
        out = subprocess.run([exe], capture_output=True, text=True)
        assert out.returncode == 0 and "bad=0" in out.stdout, out.stdout[-2000:]
