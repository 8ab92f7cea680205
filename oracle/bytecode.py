"""Oracle copy of the postfix bytecode specification (TEST INFRASTRUCTURE).

The authoritative definition is ``include/pde_b200.h``; ``tests/test_abi.py``
checks the three copies (header, product, oracle) agree.

One byte per instruction, operands folded into the byte:

  0x00            END / padding (never executed)
  0x01, 0x02      VAR0, VAR1           push coordinate (rho|r, z|x)
  0x08..0x0F      PRIM(p)              push primitive p's jet (per-problem table)
  0x10..0x13      ADD SUB MUL DIV      binary, (a b -- a op b)
  0x18..0x1B      NEG ABS SQRT EXP     unary, from infix '-', Abs(), sqrt(), exp()
  0x20..0x25      FN_NEG FN_INV FN_SQUARE FN_POW32 FN_POWN32 FN_EXPNEG
                                        the reference's six *opaque* unary names
                                        (expression_operations.py:80-89): kept
                                        distinct from NEG/... because the
                                        reference's normaliser parses them as
                                        undefined functions (lean_bridge.py:73)
  0x40..0x7F      POW(k), k<64         x ** session exponent k (constant FP64)
  0x80..0xFF      CONST(k), k<128      push session constant k (FP64)

Reserved table slots: CONST(0) = 1;  POW(0) = 3/2, POW(1) = -3/2, POW(2) = 2.
"""
from __future__ import annotations

OP_END = 0x00
OP_VAR0 = 0x01
OP_VAR1 = 0x02
OP_PRIM0 = 0x08
N_PRIM = 8
OP_ADD, OP_SUB, OP_MUL, OP_DIV = 0x10, 0x11, 0x12, 0x13
OP_NEG, OP_ABS, OP_SQRT, OP_EXP = 0x18, 0x19, 0x1A, 0x1B
OP_FN_NEG, OP_FN_INV, OP_FN_SQUARE, OP_FN_POW32, OP_FN_POWN32, OP_FN_EXPNEG = 0x20, 0x21, 0x22, 0x23, 0x24, 0x25
OP_POW0 = 0x40
N_POW = 64
OP_CONST0 = 0x80
N_CONST = 128

CONST_ONE = 0
POW_3_2, POW_N3_2, POW_2 = 0, 1, 2

# enumerator op index (oracle.enumerate triples) -> unary opcode
UNARY_OPCODES = (OP_FN_NEG, OP_FN_INV, OP_SQRT, OP_FN_SQUARE, OP_FN_POW32, OP_FN_POWN32, OP_EXP, OP_FN_EXPNEG)

FUNC_OPCODES = {
    "neg": OP_FN_NEG, "inv": OP_FN_INV, "square": OP_FN_SQUARE,
    "pow_3_2": OP_FN_POW32, "pow_neg_3_2": OP_FN_POWN32, "exp_neg": OP_FN_EXPNEG,
    "sqrt": OP_SQRT, "exp": OP_EXP, "Abs": OP_ABS,
}

# compile flags
FLAG_OK = 0
FLAG_UNSUPPORTED = 1      # token / construct the device cannot evaluate (I, zoo, x**y, ...)
FLAG_TABLE_FULL = 2       # constant or exponent table exhausted
FLAG_TOO_LONG = 4         # code longer than 255 bytes

HASH_SEED = 0x9E3779B97F4A7C15
HASH_LEN_MUL = 0xD6E8FEB86659FD93
MASK64 = (1 << 64) - 1


def mix64(x: int) -> int:
    """splitmix64 finaliser"""
    x &= MASK64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & MASK64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & MASK64
    x ^= x >> 31
    return x


def structural_hash(code: bytes) -> int:
    """64-bit structural hash of a postfix program (len + bytes)."""
    n = len(code)
    h = HASH_SEED ^ ((n * HASH_LEN_MUL) & MASK64)
    padded = code + b"\0" * ((-n) % 8)
    for i in range(0, len(padded), 8):
        w = int.from_bytes(padded[i:i + 8], "little")
        h = mix64(h ^ w)
    return h


def is_leaf(op: int) -> bool:
    return op in (OP_VAR0, OP_VAR1) or OP_PRIM0 <= op < OP_PRIM0 + N_PRIM or OP_CONST0 <= op < OP_CONST0 + N_CONST


def is_binary(op: int) -> bool:
    return OP_ADD <= op <= OP_DIV


def is_unary(op: int) -> bool:
    return (OP_NEG <= op <= OP_EXP) or (OP_FN_NEG <= op <= OP_FN_EXPNEG) or (OP_POW0 <= op < OP_POW0 + N_POW)


def stack_depth(code: bytes) -> int:
    """Maximum operand-stack depth of a postfix program; -1 if malformed."""
    sp = 0
    mx = 0
    for op in code:
        if is_leaf(op):
            sp += 1
        elif is_binary(op):
            if sp < 2:
                return -1
            sp -= 1
        elif is_unary(op):
            if sp < 1:
                return -1
        else:
            return -1
        mx = max(mx, sp)
    return mx if sp == 1 else -1
