"""Compare, symbolically, the floating-point expressions two SASS listings compute (tests/tools, no GPU needed).

Why: the literal variant (-DPM_LITERAL_NCC=1, pm_core.cuh) is meant to be BIT-identical to the reference's compiled kernels,
and under --use_fast_math that depends on which multiplies the compiler fuses into FMAs -- visible only in the SASS. This
tool executes a straight-line stretch of SASS symbolically (FMUL / FFMA / FADD / MUFU / I2F; loads are leaves), builds the
expression tree of chosen registers and prints it in a canonical form (commutative operands sorted, leaves anonymous), so
that the reference's build (cuobjdump -sass oracle/_ref/libmpmvs_ref.so) and ours can be diffed modulo register
allocation and scheduling.

    python tests/tools/sass_expr.py ref.sass  RefNccMap        # hashes of the source-fetch coordinates, weight exponents
    python tests/tools/sass_expr.py ours.sass pm_ncc_map_kernelILi0ELb0ELb0
    python tests/tools/sass_expr.py compare ref.sass BlackPixelUpdate ours.sass pm_sweep_kernelILi0ELb0ELb0
"""
import re
import sys

FP = {"FMUL": 2, "FADD": 2, "FFMA": 3}
FP2 = {"FMUL2": "FMUL", "FADD2": "FADD", "FFMA2": "FFMA"}     # sm_100 packed FP32: two lanes, each the scalar operation


def lane_operand(tok, lane, env):
    """Operand of one lane of a packed instruction: `R4.F32x2.HI_LO` (or a bare `R4`) is the register pair (R4 = lane 0, R5 = lane 1),
    `R4.F32` one register broadcast to both lanes."""
    tok = tok.strip().replace(".reuse", "")
    pair = ".F32x2" in tok or ".F32" not in tok
    tok = tok.replace(".F32x2.HI_LO", "").replace(".F32", "")
    m = re.fullmatch(r"(-?\|?)(U?R)(\d+)(\|?)", tok)
    if m and pair and lane:
        tok = f"{m.group(1)}{m.group(2)}{int(m.group(3)) + 1}{m.group(4)}"
    return operand(tok, env)


def function_body(path, name):
    out, on = [], False
    for line in open(path):
        if "Function :" in line:
            if on:
                break
            on = name in line
            continue
        if on:
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
            if m:
                out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def operand(tok, env):
    tok = tok.strip().replace(".reuse", "")
    neg = tok.startswith("-")
    tok = tok.lstrip("-")
    ab = tok.startswith("|")
    tok = tok.strip("|")
    if re.fullmatch(r"U?R\d+", tok):
        e = env.get(tok, ("leaf", tok))
    elif tok in ("RZ", "URZ"):
        e = ("const", "0")
    else:
        e = ("const", tok)
    if ab:
        e = ("abs", e)
    if neg:
        e = ("neg", e)
    return e


def canon(e):
    """Canonical text of an expression: leaves anonymous, commutative operands sorted, and negations pulled out of
    products ((-a) * b and a * (-b) round identically, so do fma(-a, b, c) and fma(a, -b, c))."""
    s, neg = _canon(e)
    return ("neg(" + s + ")") if neg else s


def _canon(e):
    k = e[0]
    if k == "leaf":
        return "x", False
    if k == "const":
        t = e[1]
        return (t[1:], True) if t.startswith("-") else (t, False)
    if k == "neg":
        s, n = _canon(e[1])
        return s, not n
    if k == "abs":
        s, _ = _canon(e[1])
        return "abs(" + s + ")", False
    if k == "i2f" and e[1][0] == "leaf":
        return "x", False       # a converted integer register is a leaf like any other (the compiler keeps some conversions in
                                # registers across code this straight-line executor cannot follow, and they then show up as leaves)
    if k in ("rcp", "sqrt", "ex2", "i2f"):
        return k + "(" + canon(e[1]) + ")", False
    if k == "FMUL":
        parts = [_canon(a) for a in e[1:]]
        return "mul(" + ",".join(sorted(p[0] for p in parts)) + ")", (sum(p[1] for p in parts) % 2 == 1)
    if k == "FADD":
        return "add(" + ",".join(sorted(canon(a) for a in e[1:])) + ")", False
    if k == "FFMA":
        parts = [_canon(a) for a in e[1:3]]
        prod = ",".join(sorted(p[0] for p in parts))
        sign = "-" if sum(p[1] for p in parts) % 2 == 1 else "+"
        return "fma(" + sign + prod + ";" + canon(e[3]) + ")", False
    return k, False


def run(body, stop_pc=None):
    """Symbolic execution in listing order (predicates and branches ignored: meant for straight-line stretches)."""
    env = {}
    for pc, ins in body:
        if stop_pc is not None and pc >= stop_pc:
            break
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        parts = ins.split(None, 1)
        op = parts[0].split(".")[0]
        args = [a.strip() for a in parts[1].split(",")] if len(parts) > 1 else []
        if op in FP and len(args) == FP[op] + 1:
            env[args[0]] = (op,) + tuple(operand(a, env) for a in args[1:])
        elif op in FP2 and len(args) == FP[FP2[op]] + 1 and re.fullmatch(r"R\d+", args[0]):
            base = int(args[0][1:])
            lanes = [(FP2[op],) + tuple(lane_operand(a, lane, env) for a in args[1:]) for lane in (0, 1)]
            env[f"R{base}"], env[f"R{base + 1}"] = lanes
        elif op == "MUFU" and len(args) == 2:
            kind = parts[0].split(".")[1].lower()
            env[args[0]] = (kind, operand(args[1], env))
        elif op == "I2FP" and len(args) == 2:
            env[args[0]] = ("i2f", operand(args[1], env))
        elif op in ("MOV", "UMOV", "R2UR") and len(args) >= 2:
            env[args[0]] = operand(args[1], env)
        elif parts[0].startswith("IMAD.MOV") and len(args) == 4 and args[1] == "RZ" and args[2] == "RZ":
            env[args[0]] = operand(args[3], env)          # IMAD.MOV.U32 Rd, RZ, RZ, Rs: a register move on the integer pipe
        elif op in ("TEX", "TLD", "TLD4") and len(args) > 1:
            for a in args[:2]:                  # TEX RZ, Rdst, ... : the fetched texel is a fresh leaf
                if re.fullmatch(r"R\d+", a):
                    env.pop(a, None)
        elif args and re.fullmatch(r"U?R\d+", args[0]):
            env.pop(args[0], None)              # anything else that writes a register: a fresh leaf
            if op in ("LDS", "LDG", "LDC", "LDCU", "LDL") and ".128" in parts[0]:
                base = int(args[0].lstrip("UR"))
                for k in range(4):
                    env.pop(f"R{base + k}", None)
            if op in ("LDS", "LDG", "LDC", "LDCU", "LDL") and ".64" in parts[0]:
                base = int(args[0].lstrip("UR"))
                for k in range(2):
                    env.pop(f"R{base + k}", None)
    return env


def main():
    import hashlib

    path, name = sys.argv[1], sys.argv[2]
    body = function_body(path, name)
    tex = [i for i, (_, ins) in enumerate(body) if ins.startswith("TEX")]
    assert tex, "no TEX in " + name
    print(f"# {name}: {len(body)} instructions, {len(tex)} texture fetches")
    # every fetch whose coordinates are fma(numerator, rcp(Z), 0.5) is a source sample of an (inlined) NCC
    for t in tex:
        env = run(body[:t])
        args = [a.strip() for a in body[t][1].split(None, 1)[1].split(",")]
        base = int(args[2].lstrip("R")) + (1 if "ARRAY" in body[t][1] else 0)     # layered fetch: the layer comes first
        cs = [canon(env.get(f"R{base + k}", ("leaf", "?"))) for k in (0, 1)]
        if not all(c.startswith("fma(") and c.endswith(";0.5)") and "rcp(" in c for c in cs):
            continue
        print(f"  TEX at {body[t][0]:#07x}: coordinates R{base},R{base + 1}  nodes {cs[0].count('(')},{cs[1].count('(')}  "
              f"sha1 {hashlib.sha1(cs[0].encode()).hexdigest()[:16]} {hashlib.sha1(cs[1].encode()).hexdigest()[:16]}")
        if len(sys.argv) > 3:
            print(cs[0])
    # the bilateral weight: argument of every MUFU.EX2 that depends on a square root (exp(-sqrt(i^2+j^2)/.. - |r-r0|/..))
    seen = {}
    for k, (pc, ins) in enumerate(body):
        if ins.startswith("MUFU.EX2"):
            env = run(body[:k])
            c = canon(operand(ins.split(",")[1], env))
            if "sqrt(" in c:
                seen.setdefault(c, []).append(pc)
    for c, pcs in seen.items():
        print(f"  EX2 argument at {[hex(p) for p in pcs][:4]}: {c}")


# The hypothesis-invariant head of ComputeHomography (cu:233-247) as the reference's SASS forms it: R_relative[k] and
# t_relative[k]. The product computes them once per problem (pm_views.h: pm_view_prep, same operations) and LOADS them, so
# in a comparison these subtrees of the reference are taken as the leaves they are in ours.
POSE_SUBTREES = ("fma(+x,x;fma(+x,x;mul(x,x)))", "fma(+add(neg(x),x),x;fma(+add(neg(x),x),x;mul(add(neg(x),x),x)))")


def fold_pose(e, memo=None):
    """The tree with every POSE_SUBTREES node replaced by a leaf."""
    memo = {} if memo is None else memo
    k = id(e)
    if k in memo:
        return memo[k]
    if e[0] in ("leaf", "const"):
        r = e
    elif e[0] in ("FFMA",) and canon(e) in POSE_SUBTREES:
        r = ("leaf", "pose")
    else:
        r = (e[0],) + tuple(fold_pose(a, memo) if isinstance(a, tuple) else a for a in e[1:])
    memo[k] = r
    return r


def source_coordinate_trees(path, name, fold=False):
    """Canonical trees of the x coordinate of every source-sample fetch (fma(numerator, rcp(Z), 0.5)) of a kernel.
    fold: take the relative pose (POSE_SUBTREES) as leaves."""
    body = function_body(path, name)
    out = []
    for t, (_, ins) in enumerate(body):
        if not ins.startswith("TEX"):
            continue
        env = run(body[:t])
        args = [a.strip() for a in ins.split(None, 1)[1].split(",")]
        base = int(args[2].lstrip("R")) + (1 if "ARRAY" in ins else 0)
        e = env.get(f"R{base}", ("leaf", "?"))
        c = canon(e)
        if c.startswith("fma(") and c.endswith(";0.5)") and "rcp(" in c:
            out.append(canon(fold_pose(e)) if fold else c)
    return out


def subtrees(s):
    stack, out = [], set()
    for i, ch in enumerate(s):
        if ch == "(":
            j = i
            while j > 0 and (s[j - 1].isalnum() or s[j - 1] == "_"):
                j -= 1
            stack.append(j)
        elif ch == ")":
            out.add(s[stack.pop():i + 1])
    return out


def compare(ref_path, ref_name, our_path, our_name):
    """Are the source coordinates the same expression in both kernels? Equal trees, or equal once ONE subtree of ours is
    taken as a leaf (a hypothesis that is still in registers in our kernel where the reference reloads it from memory)."""
    ref = source_coordinate_trees(ref_path, ref_name, fold=True)[0]
    ours = source_coordinate_trees(our_path, our_name)
    print(f"reference {ref_name}: {ref.count('(')} nodes; ours {our_name}: {len(ours)} source fetches")
    verdicts = {}
    for c in ours:
        if c == ref:
            v = "identical"
        else:
            v = "DIFFERENT"
            have = subtrees(ref)
            for k in sorted((k for k in subtrees(c) if k not in have), key=len, reverse=True):
                if c.replace(k, "x") == ref:
                    v = "identical once this is a leaf: " + (k[:120] + "..." if len(k) > 120 else k)
                    break
        verdicts[v] = verdicts.get(v, 0) + 1
    for v, n in verdicts.items():
        print(f"  {n:3d} x {v}")
    return all(not v.startswith("DIFFERENT") for v in verdicts)


PIPELINE_KERNELS = tuple("pm_exact" + k for k in ("15pm_sweep_kernel", "14pm_init_kernel", "17pm_ncc_map_kernel", "18pm_geom_map", "22pm_depth_normal", "16pm_filter"))


def pipeline_checksum(path):
    """md5 over the instruction text (addresses kept, encodings dropped) of the kernels the pipeline runs, in listing order."""
    import hashlib

    h, on = hashlib.md5(), False
    for line in open(path):
        if "Function :" in line:
            on = any(k in line for k in PIPELINE_KERNELS)
            continue
        if on and re.match(r"\s+/\*[0-9a-f]{4,6}\*/", line):
            h.update((re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line).rstrip() + "\n").encode())
    return h.hexdigest()


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "md5":
        print(pipeline_checksum(sys.argv[2]))
        sys.exit(0)
    if len(sys.argv) == 6 and sys.argv[1] == "compare":
        sys.exit(0 if compare(*sys.argv[2:6]) else 1)
    main()
