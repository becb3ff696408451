"""The arch generators of wav2letter_b200/archs.py reproduce the reference's arch files token for token (checked whenever
/root/reference is present, i.e. in the build container; the GPU box has no reference tree and uses the generators)."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("w2l_archs", os.path.join(ROOT, "wav2letter_b200", "archs.py"))
archs = importlib.util.module_from_spec(spec)
spec.loader.exec_module(archs)

REF = "/root/reference"


def tokens(text):
    return [ln.split("#")[0].split() for ln in text.splitlines() if ln.split("#")[0].split()]


@pytest.mark.parametrize("name", sorted(archs.REFERENCE_FILES))
def test_generated_arch_matches_reference_file(name):
    path = os.path.join(REF, archs.REFERENCE_FILES[name])
    if not os.path.exists(path):
        pytest.skip("reference tree not present")
    gen = archs.BASELINE_ARCHS[name][0]
    text = gen(ctc_head=False) if name == "seq2seq_tds_ctc" else gen()
    assert tokens(text) == tokens(open(path).read())


def test_ctc_head_only_changes_the_last_linear():
    a, b = tokens(archs.seq2seq_tds(True)), tokens(archs.seq2seq_tds(False))
    assert a[:-1] == b[:-1] and a[-1] == ["L", "1440", "NLABEL"] and b[-1] == ["L", "1440", "1024"]
