"""Static checks on the machine code of libfabe13.so (cuobjdump -sass; no GPU needed).

The streaming kernels must move their data with 256-bit accesses (LDG.E...256 / STG.E...256: SURVEY section 8 a-3), must
not spill on the hot path, and every kernel that produces vectors must store them as vectors.  The last check exists
because ptxas 12.9 once turned a st.global.cs.v4.b64 inside a non-inlined device function into a 64-bit store of the
first element (see the note on dense_rounds in csrc/fabe13_kernels.cu): results were wrong only on the GPU, and only on
one route."""
import re
import shutil
import subprocess

import pytest


@pytest.fixture(scope="module")
def sass(fabe13):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-sass", fabe13.build.LIBPATH], capture_output=True, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError) as e:
        pytest.skip(f"cuobjdump unavailable: {e}")
    funcs, name = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
        elif name and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            funcs[name].append(line)
    return funcs


def pick(funcs, pattern):
    hits = {n: body for n, body in funcs.items() if re.search(pattern, n)}
    assert hits, f"no kernel matches {pattern}"
    return hits


def count(body, mnemonic):
    return sum(1 for line in body if re.search(mnemonic, line))


def test_library_is_built_for_sm_100a(fabe13):
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-lelf", fabe13.build.LIBPATH], capture_output=True, text=True, check=True).stdout
    except (OSError, subprocess.CalledProcessError) as e:
        pytest.skip(f"cuobjdump unavailable: {e}")
    assert "sm_100a" in out, out


# template arguments in the mangled names: <CORE, VEC, UNROLL, WITH_K, OUT[, ONE_TILE]>
@pytest.mark.parametrize("core", [0, 1])
def test_default_hot_kernel_uses_256_bit_accesses_and_does_not_spill(sass, core):
    (name, body), = pick(sass, rf"sincos_fused_kernelILi{core}ELi4ELi1ELb0ELi0ELb1E").items()
    assert count(body, r"LDG\.E\.NA\S*\.256") == 1, name          # one vector load per thread and tile
    assert count(body, r"STG\.E\.EF\S*\.256") == 2, name          # sines, cosines
    # no spills on the streaming path; the out-of-line dense routine parks three words once per vector (see
    # test_dense_route_uses_the_fast_reduction_and_keeps_the_exact_one_out_of_line)
    first_call_target = min([int(m.group(1), 16) for line in body for m in [re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", line)] if m])
    main = [line for line in body if (m := re.search(r"/\*([0-9a-f]{4,6})\*/", line)) and int(m.group(1), 16) < first_call_target]
    assert len(main) > 400 and count(main, r"\b(STL|LDL)\b") == 0, f"{name} spills registers on the streaming path"
    assert count(body, r"\bSTL\b") <= 6 and count(body, r"\bLDL\b") <= 3, f"{name} spills registers"
    assert count(body, r"\bBAR\b") == 0, f"{name}: the hot path must not contain a block-wide barrier"
    assert count(body, r"\bDFMA\b") >= 18                         # the FP64 work is there (not optimised away)


def test_every_vec4_streaming_kernel_stores_vectors(sass):
    for name, body in pick(sass, r"sincos_(fused|stream)_kernelILi[01]ELi4E").items():
        m = re.search(r"kernelILi[01]ELi4ELi([12])ELb[01]ELi([0-5])E", name)
        unroll, out = int(m.group(1)), int(m.group(2))
        per_vector = 2 if out == 0 else 1
        assert count(body, r"STG\.E\.EF\S*\.256") == per_vector * unroll, name
        assert count(body, r"LDG\.E\.NA\S*\.256") == unroll, name
        assert count(body, r"STG\.E\.EF\.64\b") == 0, f"{name}: a vector store was narrowed"


def test_narrower_vector_widths_exist_for_incongruent_pointers(sass):
    assert pick(sass, r"sincos_fused_kernelILi0ELi2E")
    assert pick(sass, r"sincos_fused_kernelILi0ELi1E")
    for name, body in pick(sass, r"sincos_fused_kernelILi0ELi2ELi1ELb0ELi0E").items():
        assert count(body, r"LDG\.E\.NA\S*\.128") == 1 and count(body, r"STG\.E\.EF\S*\.128") == 2, name


def test_fp32_and_derived_kernels_keep_vector_accesses_and_do_not_spill(sass):
    for name, body in pick(sass, r"sincosf_kernelILi0E").items():
        assert count(body, r"LDG\.E\.NA\S*\.256") == 1 and count(body, r"STG\.E\.EF\S*\.256") == 2, name   # eight floats per thread
        assert count(body, r"\b(STL|LDL)\b") == 0, f"{name} spills registers"
        assert count(body, r"\bF2F\.F32\.F64\b") >= 16                                                       # two roundings per element
    for out in (3, 4, 5):  # tan, cot, sinc: one output vector, the checks-free division (MUFU.RCP64H) inline
        (name, body), = pick(sass, rf"sincos_fused_kernelILi0ELi4ELi1ELb0ELi{out}ELb1E").items()
        assert count(body, r"STG\.E\.EF\S*\.256") == 1 and count(body, r"\b(STL|LDL)\b") == 0, name
        assert count(body, r"MUFU\.RCP64H") >= 4, name


def _hot_kernel_functions(sass):
    """The default hot kernel's SASS split at its out-of-line routines: {call target address: lines from there to the next
    RET} plus the kernel text itself."""
    (name, body), = pick(sass, r"sincos_fused_kernelILi0ELi4ELi1ELb0ELi0ELb1E").items()
    text = [re.sub(r"/\* 0x[0-9a-f]+ \*/", "", line) for line in body]
    targets = sorted({int(m.group(1), 16) for line in text for m in [re.search(r"CALL\.REL\.NOINC (0x[0-9a-f]+)", line)] if m})
    funcs = {}
    for t in targets:
        start = next(i for i, line in enumerate(text) if f"/*{t:04x}*/" in line)
        end = next(i for i in range(start, len(text)) if re.search(r"\bRET\.REL", text[i]))
        funcs[t] = text[start:end + 1]
    return name, text, funcs


def test_dense_route_uses_the_fast_reduction_and_keeps_the_exact_one_out_of_line(sass):
    """Arguments >= 2^45 are reduced by the fast form (csrc/fabe13_core.h 2c): one 256-bit read-only, evict-last load of
    the lane's table entry and FP64 arithmetic, no integer product.  Checked on the dense-warp routine of the default hot
    kernel (dense_vector: out of line, one vector in and out in registers, four unrolled rounds): four such loads, no
    IMAD.WIDE, no shared memory at all, one override vote per round, and short -- the rolled loop of round 2's first
    version ran 138 instructions per element with the 192-bit product inline, 196 before that product was rewritten."""
    name, text, funcs = _hot_kernel_functions(sass)
    dense = [f for f in funcs.values() if count(f, r"LDG\.E\S*\.256\.CONSTANT") >= 4]
    assert len(dense) == 1, name
    dense = dense[0]
    assert count(dense, r"LDG\.E\.EL\S*\.256\.CONSTANT") == 4, name   # the table entry: one load per element, kept in L1
    assert count(dense, r"IMAD\.WIDE") == 0, name                       # no integer product on the hot path
    assert count(dense, r"\b(LDS|STS)\b") == 0, name                    # vectors stay in registers
    assert count(dense, r"\b(DFMA|DMUL|DADD)\b") >= 160, name           # 4 x (reduction 20 + core 15 + the other range's k, r)
    assert count(dense, r"\bVOTE\.ANY\b") == 4, name                   # the overrides sit behind a warp vote per round ...
    assert count(dense, r"CALL\.REL") == 8, name                        # ... and real branches around two calls per round
    assert count(dense, r"\bSTL\b") <= 6 and count(dense, r"\bLDL\b") <= 3, name   # (three words parked once per vector)
    assert len(dense) <= 520, len(dense)                                 # <= 130 per element


def test_payne_hanek_product_has_no_accumulator_moves(sass):
    """The 53 x 192-bit product of the EXACT Payne-Hanek reduction (csrc/fabe13_core.h 2b; out of line since the fast form
    took over the hot path) is nine independent IMAD.WIDE with a zero addend plus carry chains.  Its first form (a running
    64-bit accumulator per chain) needed two register moves per step to build the next addend pair, which showed up as
    `VIADD Rx, Ry, URz` with a zeroed uniform register and cost 7 % on the extreme-range workload
    (profiles/r01_ab_ph_multiply.txt)."""
    name, text, funcs = _hot_kernel_functions(sass)
    exact = [f for f in funcs.values() if count(f, r"IMAD\.WIDE\.U32") >= 7]
    assert len(exact) == 1, name
    exact = exact[0]
    wide = [line for line in exact if "IMAD.WIDE.U32" in line]
    assert 7 <= len(wide) <= 9, wide                               # (two of the nine may be emitted as IMAD.HI / IMAD)
    assert all(re.search(r", RZ ;", line) for line in wide), wide   # independent products: no 64-bit addend
    assert count(exact, r"VIADD R\d+, R\d+, UR\d+ ;") == 0, name
    assert count(exact, r"\b(STL|LDL)\b") == 0, name
    assert len(exact) <= 110, len(exact)
