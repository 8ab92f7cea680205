#!/usr/bin/env python3
"""Micro-op statistics of the synthetic depth-5 mix: a Python mirror of translate()
(pde_engine_b200/csrc/validate.cuh) run over oracle.synth trees.  Development tool."""
import collections, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import bytecode as bc, synth

T, S = "T", "S"
def is_prim(b): return bc.OP_PRIM0 <= b < bc.OP_PRIM0 + bc.N_PRIM
def leaf_kind(b): return "C" if b >= bc.OP_CONST0 else "P" if is_prim(b) else "V"

def translate(code):
    uc = []; vst = []; tpos = [-1]
    def spill():
        if tpos[0] >= 0: uc.append("SPILL"); vst[tpos[0]] = S
    def set_leaf(b): uc.append("SET" + leaf_kind(b))
    def blr(o, leaf):
        k = leaf_kind(leaf)
        if k == "C": uc.append(["ADDC", "SUBC", "MULC", "MULRC"][o])
        elif k == "V": uc.append(["ADDV", "SUBV", "MULV", "DIVV"][o])
        else: uc.append(["ADD_P", "SUB_P", "MUL_P", "DIV_P"][o])
    for b in code:
        if bc.is_leaf(b): vst.append(b)
        elif bc.is_unary(b):
            top = vst[-1]
            if top != T:
                spill(); set_leaf(top); vst[-1] = T; tpos[0] = len(vst) - 1
            uc.append({bc.OP_FN_NEG: "NEG", bc.OP_NEG: "NEG", bc.OP_ABS: "ABS", bc.OP_SQRT: "SQRT", bc.OP_EXP: "EXP",
                       bc.OP_FN_INV: "INV", bc.OP_FN_SQUARE: "SQUARE", bc.OP_FN_POW32: "POW", bc.OP_FN_POWN32: "POW",
                       bc.OP_FN_EXPNEG: "EXPN"}.get(b, "POW"))
        else:
            bb = vst.pop(); aa = vst.pop(); o = b - bc.OP_ADD
            if aa == S and bb == T:
                uc.append(["ADD_S", "RSUB_S", "MUL_S", "RDIV_S"][o])
            elif aa == T and bb != S: blr(o, bb)
            elif bb == T and aa != S:
                if o in (0, 2): blr(o, aa)
                elif o == 1:
                    if aa >= bc.OP_CONST0: uc.append("RSUBC")
                    else: uc.append("NEG"); blr(0, aa)
                elif is_prim(aa): uc.append("RDIV_P")
                else: uc.append("INV"); blr(2, aa)
            else:
                spill(); set_leaf(aa); blr(o, bb)
            vst.append(T); tpos[0] = len(vst) - 1
    if vst[0] != T: set_leaf(vst[0])
    uc.append("END")
    return uc

if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    uni = collections.Counter(); bi = collections.Counter(); tot = 0
    for c in synth.trees(synth.SEED_TREES, 0, n, 5):
        u = translate(c); tot += len(u)
        uni.update(u); bi.update(zip(u, u[1:]))
    print(f"micro-ops per tree: {tot / n:.3f}")
    for k, v in uni.most_common(): print(f"  {k:8s} {v / n:.3f}")
    print("bigrams:")
    for k, v in bi.most_common(25): print(f"  {k[0]:8s} {k[1]:8s} {v / n:.3f}")
