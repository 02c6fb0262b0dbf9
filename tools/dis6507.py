#!/usr/bin/env python3
"""Linear/recursive 6507 disassembler for a 2 KiB Atari 2600 cart (development aid).

Usage: python tools/dis6507.py ROM.bin [start_hex end_hex]
"""
import sys

# mnemonic, addressing mode, cycles
M = {}
def op(code, mn, mode, cyc): M[code] = (mn, mode, cyc)
for mn, base in (("ORA",0x00),("AND",0x20),("EOR",0x40),("ADC",0x60),("STA",0x80),("LDA",0xA0),("CMP",0xC0),("SBC",0xE0)):
    op(base+0x01, mn, "izx", 6); op(base+0x05, mn, "zp", 3); op(base+0x09, mn, "imm", 2); op(base+0x0D, mn, "abs", 4)
    op(base+0x11, mn, "izy", 5); op(base+0x15, mn, "zpx", 4); op(base+0x19, mn, "aby", 4); op(base+0x1D, mn, "abx", 4)
del M[0x89]
for mn, base in (("ASL",0x00),("ROL",0x20),("LSR",0x40),("ROR",0x60)):
    op(base+0x06, mn, "zp", 5); op(base+0x0A, mn, "acc", 2); op(base+0x0E, mn, "abs", 6); op(base+0x16, mn, "zpx", 6); op(base+0x1E, mn, "abx", 7)
for mn, base in (("DEC",0xC0),("INC",0xE0)):
    op(base+0x06, mn, "zp", 5); op(base+0x0E, mn, "abs", 6); op(base+0x16, mn, "zpx", 6); op(base+0x1E, mn, "abx", 7)
op(0x86,"STX","zp",3); op(0x96,"STX","zpy",4); op(0x8E,"STX","abs",4)
op(0x84,"STY","zp",3); op(0x94,"STY","zpx",4); op(0x8C,"STY","abs",4)
op(0xA2,"LDX","imm",2); op(0xA6,"LDX","zp",3); op(0xB6,"LDX","zpy",4); op(0xAE,"LDX","abs",4); op(0xBE,"LDX","aby",4)
op(0xA0,"LDY","imm",2); op(0xA4,"LDY","zp",3); op(0xB4,"LDY","zpx",4); op(0xAC,"LDY","abs",4); op(0xBC,"LDY","abx",4)
op(0xE0,"CPX","imm",2); op(0xE4,"CPX","zp",3); op(0xEC,"CPX","abs",4)
op(0xC0,"CPY","imm",2); op(0xC4,"CPY","zp",3); op(0xCC,"CPY","abs",4)
op(0x24,"BIT","zp",3); op(0x2C,"BIT","abs",4)
for code, mn in ((0x10,"BPL"),(0x30,"BMI"),(0x50,"BVC"),(0x70,"BVS"),(0x90,"BCC"),(0xB0,"BCS"),(0xD0,"BNE"),(0xF0,"BEQ")):
    op(code, mn, "rel", 2)
for code, mn, cyc in ((0x00,"BRK",7),(0x08,"PHP",3),(0x18,"CLC",2),(0x28,"PLP",4),(0x38,"SEC",2),(0x40,"RTI",6),(0x48,"PHA",3),
                      (0x58,"CLI",2),(0x60,"RTS",6),(0x68,"PLA",4),(0x78,"SEI",2),(0x88,"DEY",2),(0x8A,"TXA",2),(0x98,"TYA",2),
                      (0x9A,"TXS",2),(0xA8,"TAY",2),(0xAA,"TAX",2),(0xB8,"CLV",2),(0xBA,"TSX",2),(0xC8,"INY",2),(0xCA,"DEX",2),
                      (0xD8,"CLD",2),(0xE8,"INX",2),(0xEA,"NOP",2),(0xF8,"SED",2)):
    op(code, mn, "imp", cyc)
op(0x20,"JSR","abs",6); op(0x4C,"JMP","abs",3); op(0x6C,"JMP","ind",5)
LEN = {"imp":1,"acc":1,"imm":2,"zp":2,"zpx":2,"zpy":2,"izx":2,"izy":2,"rel":2,"abs":3,"abx":3,"aby":3,"ind":3}

TIAW = {0:"VSYNC",1:"VBLANK",2:"WSYNC",3:"RSYNC",4:"NUSIZ0",5:"NUSIZ1",6:"COLUP0",7:"COLUP1",8:"COLUPF",9:"COLUBK",10:"CTRLPF",
        11:"REFP0",12:"REFP1",13:"PF0",14:"PF1",15:"PF2",16:"RESP0",17:"RESP1",18:"RESM0",19:"RESM1",20:"RESBL",21:"AUDC0",22:"AUDC1",
        23:"AUDF0",24:"AUDF1",25:"AUDV0",26:"AUDV1",27:"GRP0",28:"GRP1",29:"ENAM0",30:"ENAM1",31:"ENABL",32:"HMP0",33:"HMP1",34:"HMM0",
        35:"HMM1",36:"HMBL",37:"VDELP0",38:"VDELP1",39:"VDELBL",40:"RESMP0",41:"RESMP1",42:"HMOVE",43:"HMCLR",44:"CXCLR"}
TIAR = {0x30:"CXM0P",0x31:"CXM1P",0x32:"CXP0FB",0x33:"CXP1FB",0x34:"CXM0FB",0x35:"CXM1FB",0x36:"CXBLPF",0x37:"CXPPMM",
        0x38:"INPT0",0x39:"INPT1",0x3A:"INPT2",0x3B:"INPT3",0x3C:"INPT4",0x3D:"INPT5"}
RIOT = {0x280:"SWCHA",0x281:"SWACNT",0x282:"SWCHB",0x283:"SWBCNT",0x284:"INTIM",0x294:"TIM1T",0x295:"TIM8T",0x296:"TIM64T",0x297:"T1024T"}

def name(a, write):
    if a < 0x80:
        if write and (a & 0x3f) in TIAW: return TIAW[a & 0x3f]
        if not write and (a | 0x30) in TIAR and a >= 0x30: return TIAR[a]
    if a in RIOT: return RIOT[a]
    return None

def fmt(rom, pc):
    b = rom[pc & 0x7ff]
    if b not in M:
        return 1, ".byte $%02X" % b, None
    mn, mode, cyc = M[b]
    n = LEN[mode]
    lo = rom[(pc+1) & 0x7ff] if n > 1 else 0
    hi = rom[(pc+2) & 0x7ff] if n > 2 else 0
    w = mn in ("STA","STX","STY")
    if mode in ("imp",): s = mn
    elif mode == "acc": s = mn + " A"
    elif mode == "imm": s = "%s #$%02X" % (mn, lo)
    elif mode == "zp": s = "%s $%02X" % (mn, lo)
    elif mode == "zpx": s = "%s $%02X,X" % (mn, lo)
    elif mode == "zpy": s = "%s $%02X,Y" % (mn, lo)
    elif mode == "izx": s = "%s ($%02X,X)" % (mn, lo)
    elif mode == "izy": s = "%s ($%02X),Y" % (mn, lo)
    elif mode == "rel":
        t = (pc + 2 + (lo - 256 if lo > 127 else lo)) & 0xffff
        s = "%s $%04X" % (mn, t)
    elif mode == "abs": s = "%s $%04X" % (mn, lo | hi << 8)
    elif mode == "abx": s = "%s $%04X,X" % (mn, lo | hi << 8)
    elif mode == "aby": s = "%s $%04X,Y" % (mn, lo | hi << 8)
    elif mode == "ind": s = "%s ($%04X)" % (mn, lo | hi << 8)
    a = lo | hi << 8 if n == 3 else lo
    if mode in ("zp","zpx","zpy","abs","abx","aby") and mn not in ("JMP","JSR"):
        nm = name(a, w)
        if nm: s += "   ; " + nm
    return n, s, cyc

if __name__ == "__main__":
    rom = open(sys.argv[1], "rb").read()
    start = int(sys.argv[2], 16) if len(sys.argv) > 2 else 0xF000
    end = int(sys.argv[3], 16) if len(sys.argv) > 3 else 0xF658
    pc = start
    while pc < end:
        n, s, cyc = fmt(rom, pc)
        raw = " ".join("%02X" % rom[(pc+i) & 0x7ff] for i in range(n))
        print("%04X  %-9s %-28s %s" % (pc, raw, s, cyc if cyc else ""))
        pc += n
