"""What the shipped cubin contains, read with cuobjdump (no GPU needed): sm_100a code only, the candidate loop of the
importance-sampler scoring kernel with exactly the instruction mix DESIGN.md 4.1 states, no local-memory traffic inside any
candidate loop, the programmatic-dependent-launch instructions of the kernel chain, and REDUX only in the low-K instantiation."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "compression_without_quantization_b200", "libcwq.so")
CUOBJDUMP = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"

pytestmark = pytest.mark.skipif(not (os.path.exists(LIB) and os.path.exists(CUOBJDUMP)),
                                reason="needs the built libcwq.so and cuobjdump")


@pytest.fixture(scope="module")
def sass():
    """function name -> list of (address, instruction text)"""
    out = subprocess.run([CUOBJDUMP, "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = funcs.setdefault(m.group(1), [])
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur is not None:
            cur.append((int(m.group(1), 16), re.sub(r"^@!?U?P\w+\s+", "", m.group(2).strip())))
    return funcs, out


def _loops(ins):
    """bodies of the innermost-looking loops: [target, branch] of every backward branch"""
    for a, t in ins:
        m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
        if m and int(m.group(1), 16) < a:
            tgt = int(m.group(1), 16)
            yield [x[1] for x in ins if tgt <= x[0] <= a]


def _find(funcs, *parts):
    names = [n for n in funcs if all(p in n for p in parts)]
    assert len(names) == 1, names
    return funcs[names[0]]


def test_only_sm_100a_code(sass):
    _, text = sass
    archs = set(re.findall(r"arch = (sm_\w+)", text))
    assert archs == {"sm_100a"}, archs


def test_candidate_loop_instruction_mix(sass):
    """The 16-dim candidate loop of the default k_is_score instantiation (SMALL = 1, LOWK = 0): 231 instructions per 16
    candidate-dims -- 44 wide multiplies (3 Philox blocks x 14 + 2 shared), 32 MUFU (8 Box-Muller pairs x lg2, sqrt, cos,
    sin), 66 LOP3 + 8 PRMT, 73 FP32 -- and nothing else: no loads from local memory, no shuffles, no barriers."""
    funcs, _ = sass
    ins = _find(funcs, "k_is_score", "ILb1ELb0E")
    full = [b for b in _loops(ins) if sum("MUFU" in x for x in b) == 32 and len(b) < 300]
    assert len(full) == 1
    ops = collections.Counter(x.split()[0].split(".")[0] for x in full[0])
    assert len(full[0]) == 231
    assert ops["IMAD"] == 44 and all("IMAD.WIDE" in x for x in full[0] if x.startswith("IMAD"))
    assert ops["MUFU"] == 32 and ops["LOP3"] == 66 and ops["PRMT"] == 8
    assert ops["FFMA"] + ops["FMUL"] + ops["FADD"] == 73
    assert not any(x.split()[0].split(".")[0] in ("LDL", "STL", "SHFL", "BAR", "LDG", "STG", "ATOM", "ATOMG", "RED") for x in full[0])


def test_greedy_candidate_loop_instruction_mix(sass):
    """One 16-dim unit of one candidate in both greedy kernels: the same Philox / Box-Muller work as the
    importance sampler's loop plus 11 shared-memory loads (8 of (c, e) pairs, 3 of block constants) and 2 FFMA per dim."""
    funcs, _ = sass
    for kernel in ("k_greedy_encode", "k_greedy_lockstep"):
        ins = _find(funcs, kernel)
        unit = [b for b in _loops(ins) if sum("MUFU" in x for x in b) == 32 and len(b) < 300]
        assert len(unit) == 1, kernel
        ops = collections.Counter(x.split()[0].split(".")[0] for x in unit[0])
        assert ops["MUFU"] == 32 and ops["LDS"] == 11 and ops["LOP3"] + ops["PRMT"] == 72, (kernel, ops)
        assert 44 <= ops["IMAD"] <= 46 and ops["FFMA"] + ops["FMUL"] + ops["FADD"] == 72, (kernel, ops)
        assert len(unit[0]) <= 240, (kernel, len(unit[0]))
        assert not any(x.startswith(("LDL", "STL")) for x in unit[0]), kernel


def test_no_local_memory_in_any_candidate_loop(sass):
    funcs, _ = sass
    for name, ins in funcs.items():
        if "k_is_score" not in name and "k_greedy" not in name:
            continue
        for body in _loops(ins):
            n_mufu = sum("MUFU" in x for x in body)
            if n_mufu >= 4 and n_mufu % 4 == 0 and len(body) < 400:   # a candidate loop (whole Box-Muller pairs inside), not an outer loop
                assert not any(x.startswith(("LDL", "STL")) for x in body), name


def test_half_quad_variants_exist(sass):
    """One loop per Box-Muller pair count 1 .. 8 of a <= 16-dim group: 4 MUFU per pair."""
    funcs, _ = sass
    ins = _find(funcs, "k_is_score", "ILb1ELb0E")
    mufu = sorted(sum("MUFU" in x for x in b) for b in _loops(ins) if 4 <= sum("MUFU" in x for x in b) <= 32 and len(b) < 300)
    assert mufu == [4, 8, 12, 16, 20, 24, 28, 32], mufu


def test_dependent_launch_chain_and_redux(sass):
    """griddepcontrol.wait is ACQBULK and griddepcontrol.launch_dependents is PREEXIT in sm_100a SASS: every kernel of the
    fused pipeline's chain waits before it does anything, all but the scoring kernel let their successor in early; REDUX (the
    two-instruction warp argmax) belongs to the low-K instantiations of the scoring kernel only."""
    funcs, _ = sass
    chain = ["k_front_tiles", "k_front_chain", "k_front_groups", "k_front_prefix", "k_is_index", "k_finish_dims", "k_pack_elias",
             "k_bits_to_ascii"]
    seen = 0
    for name, ins in funcs.items():
        ops = [t.split()[0].split(".")[0] for _, t in ins]
        if any(k in name for k in chain):
            seen += 1
            assert "ACQBULK" in ops and "PREEXIT" in ops, name
            assert ops.index("ACQBULK") < 40, name          # the wait sits at the top, before any real work
        if "k_is_score" in name:
            seen += 1
            assert "ACQBULK" in ops and "PREEXIT" not in ops, name
            low_k = "Lb1EEEv" in name                         # k_is_score<SMALL, LOWK = true>
            assert any(o.endswith("REDUX") for o in ops) == low_k, name      # CREDUX.MAX / CREDUX.MIN on sm_100a
    assert seen == len(chain) + 4
