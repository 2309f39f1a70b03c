"""CPU: the C-ABI library loads and exports every symbol include/abides_b200.h declares; struct
mirrors agree with the header's sizes.  No compute calls (there is no GPU here)."""
import ctypes
import os
import re
import subprocess

from abides_b200 import _abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_declares_what_the_library_exports():
    hdr = open(os.path.join(ROOT, "include", "abides_b200.h")).read()
    declared = set(re.findall(r"\b(abides_[a-z_]+)\s*\(", hdr))
    assert declared == set(_abi.EXPORTS)
    L = _abi.lib()
    for name in declared:
        assert hasattr(L, name), name


def test_struct_sizes_match_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "abides_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu\\n",'
                   "sizeof(abides_agent_type),sizeof(abides_instance_cfg),sizeof(abides_batch),"
                   "sizeof(abides_instance_result),sizeof(abides_agent_result),sizeof(abides_trace_rec),"
                   "sizeof(abides_book_op),sizeof(abides_book_event),sizeof(abides_book_state));return 0;}\n")
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(t) for t in (_abi.AgentType, _abi.InstanceCfg, _abi.Batch, _abi.InstanceResult,
                                        _abi.AgentResult, _abi.TraceRec, _abi.BookOp, _abi.BookEvent,
                                        _abi.BookState)]
    assert got == want


def test_create_fails_loudly_without_a_device():
    import torch
    if torch.cuda.is_available():
        return
    from abides_b200.engine import Engine, EngineError
    try:
        Engine(0)
    except EngineError as e:
        assert "no CUDA device" in str(e) or "CUDA" in str(e)
    else:
        raise AssertionError("Engine() must not succeed without a GPU")
