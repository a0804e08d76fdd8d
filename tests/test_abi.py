"""CPU tier: the C-ABI library loads, exports every symbol include/rxm.h declares,
and its host-only entry points (table text, validation) work.  No compute calls."""
import ctypes as C
import os
import re

import pytest

import helpers as H
from cases import CASE_NAMES, load_case

rxm = H.rxm


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(H.ROOT, "include", "rxm.h")).read()
    declared = set(re.findall(r"\b(rxm_[a-z_]+)\s*\(", header))
    declared -= {"rxm_handle"}
    assert declared == set(rxm.ABI_SYMBOLS), declared ^ set(rxm.ABI_SYMBOLS)
    L = C.CDLL(rxm.LIB_PATH)
    for name in declared:
        assert getattr(L, name) is not None


@pytest.mark.parametrize("name", CASE_NAMES)
def test_table_text_roundtrip(name):
    t, _, _ = load_case(name)
    text = t.format()
    assert text == t.text
    t2 = rxm.Tables(text)
    assert t2.c.n_states == t.c.n_states and t2.c.n_edges == t.c.n_edges
    assert rxm.lib().rxm_tables_validate(t2.ptr) == rxm.RXM_OK


@pytest.mark.parametrize("bad", [
    "", "rxm-tables 2\n", "rxm-tables 1\nkind dfa\n",
    "rxm-tables 1\nkind nfa\nreversed 0\nstates 2\nstart 0\nfinish 5\ncells 0\nedges 0\nend\n",
    "rxm-tables 1\nkind nfa\nreversed 0\nstates 2\nstart 0\nfinish 1\ncells 0\nedges 1\n0 L a 7\nend\n",
    "rxm-tables 1\nkind nfa\nreversed 0\nstates 2\nstart 0\nfinish 1\ncells 0\nedges 1\n0 L a 1 o1\nend\n",
    "rxm-tables 1\nkind mfa\nreversed 0\nstates 2\nstart 0\nfinish 1\ncells 1\nedges 1\n0 L a 1 o1 c1\nend\n",
    "rxm-tables 1\nkind nfa\nreversed 0\nstates 2\nstart 0\nfinish 1\ncells 0\nedges 2\n0 L a 1\nend\n",
])
def test_table_parse_rejects_malformed(bad):
    with pytest.raises(rxm.RxmError) as e:
        rxm.Tables(bad)
    assert e.value.status == rxm.RXM_ERR_PARSE


def test_strerror_and_null_arguments():
    assert rxm.strerror(rxm.RXM_OK) == "ok"
    assert "fallback" in rxm.strerror(rxm.RXM_ERR_UNSUPPORTED)
    L = rxm.lib()
    assert L.rxm_tables_validate(None) == rxm.RXM_ERR_INVALID
    assert L.rxm_match_batch(None, None, None, 0, None, None) == rxm.RXM_ERR_INVALID
    assert L.rxm_match_text(None, None, 0, None, 0, None, None, None) == rxm.RXM_ERR_INVALID
    assert L.rxm_free(None) == rxm.RXM_OK


def test_no_device_is_an_error_not_a_fallback():
    """Without a CUDA device the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    t, _, _ = load_case("nfa_config2")
    with pytest.raises(rxm.RxmError) as e:
        rxm.Matcher(t, 0)
    assert e.value.status in (rxm.RXM_ERR_NO_DEVICE, rxm.RXM_ERR_CUDA)


# ---- planner limits: reported as RXM_ERR_UNSUPPORTED at upload, never a fallback ------------
EPS_CYCLE_NFA = """rxm-tables 1
kind nfa
reversed 0
states 3
start 0
finish 2
cells 0
edges 3
0 E - 1
1 E - 0
1 L a 2
end
"""

EPS_CYCLE_MFA = """rxm-tables 1
kind mfa
reversed 0
states 3
start 0
finish 2
cells 1
edges 3
0 E - 1
1 E - 0
1 L a 2 o1
end
"""


@pytest.mark.gpu
@pytest.mark.parametrize("text", [EPS_CYCLE_NFA, EPS_CYCLE_MFA])
def test_epsilon_cycle_is_unsupported(text):
    """The reference recurses without bound on an epsilon cycle (automata.cpp:108-110,
    mfa.cpp:143-147: stack overflow); the planner refuses the table instead."""
    t = rxm.Tables(text)
    with pytest.raises(rxm.RxmError) as e:
        rxm.Matcher(t, 0)
    assert e.value.status == rxm.RXM_ERR_UNSUPPORTED


def test_every_kernel_launch_goes_through_rxm_launch():
    """The CPU tier runs the kernel sources on the SIMT emulator (tests/hostsim) only as long as
    no launch bypasses RXM_LAUNCH and no kernel declares its dynamic shared memory by hand."""
    import glob
    import re
    csrc = os.path.join(H.PKG, "csrc")
    for path in glob.glob(os.path.join(csrc, "*.cu")):
        text = open(path).read()
        text = re.sub(r"//[^\n]*", "", text)
        assert "<<<" not in text, path
        assert "extern __shared__" not in text, path
    header = re.sub(r"//[^\n]*", "", open(os.path.join(csrc, "rxm_kernels.cuh")).read())
    assert header.count("<<<") == 1 and "RXM_LAUNCH" in header
