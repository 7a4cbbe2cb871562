"""CPU checks of the C-ABI boundary: the library builds/loads, exports every symbol the header
declares, the ctypes structs have the sizes the C compiler gives them, and argument validation
rejects bad calls without touching a GPU."""
import ctypes as C
import re
import subprocess
import tempfile
from pathlib import Path

import pytest

from sac_td3_cudagraphs_pytorch_b200 import _lib as L

REPO = Path(__file__).resolve().parent.parent
HEADER = REPO / "include" / "b2rl.h"


def declared_functions():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(b2rl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = L.load()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/b2rl.h but not exported"
        assert n in L.SYMBOLS, f"{n} has no ctypes prototype"
    assert lib.b2rl_version() == 111


def test_prototypes_have_the_declared_number_of_arguments():
    """Every ctypes prototype in _lib.SYMBOLS lists as many arguments as the declaration in include/b2rl.h."""
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    seen = 0
    for m in re.finditer(r"\b(b2rl_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert len(L.SYMBOLS[name][1]) == n, f"{name}: header declares {n} arguments, binding has {len(L.SYMBOLS[name][1])}"
        seen += 1
    assert seen == len(L.SYMBOLS)


def test_struct_layouts_match_the_c_compiler():
    src = '#include <stdio.h>\n#include "b2rl.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",' \
          'sizeof(b2rl_net_t),sizeof(b2rl_rowfmt_t),sizeof(b2rl_hyper_t),sizeof(b2rl_update_args_t),' \
          'sizeof(b2rl_seg_t),sizeof(b2rl_adam_args_t),__builtin_offsetof(b2rl_update_args_t, workspace),' \
          'sizeof(b2rl_stack_t),sizeof(b2rl_wide_policy_t),sizeof(b2rl_wide_q_t),__builtin_offsetof(b2rl_stack_t, out_stride),' \
          'sizeof(b2rl_colsum_job_t));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        c = Path(d) / "s.c"
        c.write_text(src)
        subprocess.check_call(["gcc", "-I", str(REPO / "include"), str(c), "-o", str(Path(d) / "s")])
        out = subprocess.check_output([str(Path(d) / "s")]).split()
    got = [C.sizeof(L.Net), C.sizeof(L.RowFmt), C.sizeof(L.Hyper), C.sizeof(L.UpdateArgs), C.sizeof(L.Seg),
           C.sizeof(L.AdamArgs), L.UpdateArgs.workspace.offset, C.sizeof(L.Stack), C.sizeof(L.WidePolicy), C.sizeof(L.WideQ),
           L.Stack.out_stride.offset, C.sizeof(L.ColsumJob)]
    assert [int(x) for x in out] == got


def test_bad_arguments_are_rejected_before_any_launch():
    lib = L.load()
    assert lib.b2rl_workspace_floats(0) == -1
    assert lib.b2rl_workspace_floats(6) > 0  # ragged batches are legal: the tail of the last 16-row group is masked
    assert lib.b2rl_workspace_floats(256) > 0
    fmt = L.RowFmt(11, 3, 27, 0)  # row_stride not a multiple of 4
    rc = lib.b2rl_replay_sample_gather(16, 0, 10, fmt, 4, 1, None, None, 16, 0, None, 3, 1, 0, None)
    assert rc == -1 and b"row_stride" in lib.b2rl_last_error()
    a = L.UpdateArgs()
    assert lib.b2rl_critic_update_sac(C.byref(a), None) == -1
    ad = L.AdamArgs()
    assert lib.b2rl_adam_polyak_multi(C.byref(ad), None) == -1
    with pytest.raises(L.B2rlError):
        L.check(-1, "x")


def test_layout_is_aligned_and_twin_critics_are_congruent():
    from sac_td3_cudagraphs_pytorch_b200.arena import make_layout
    for ob, ac, td3, ln in [(11, 3, False, True), (376, 17, False, True), (4, 1, True, False), (17, 6, True, True)]:
        lay = make_layout(ob, ac, td3, ln)
        c0, c1 = lay.critic
        assert lay.region % 4 == 0
        for net in (c0, c1, lay.actor):
            for f, o in net.off.items():
                assert o == -1 or o % 4 == 0
            assert net.begin % 4 == 0 and net.end % 4 == 0 and net.core_end <= net.off["w2n"]
        assert {f: o - c0.begin for f, o in c0.off.items() if o >= 0} == \
               {f: o - c1.begin for f, o in c1.off.items() if o >= 0}
        assert lay.actor.out_dim == (ac if td3 else 2 * ac)
        n_params = sum(net.numel(f) for net in (c0,) for f, o in net.off.items() if o >= 0 and f != "w2n")
        assert n_params == (ob + ac) * 256 + 256 * 256 + 256 + 2 * 256 + (4 * 256 if ln else 0) + 1
