"""Token / target pipeline (host code in wav2letter_b200/host/text_pipeline.cpp through the C ABI):
the dataset's target transform and evalOutput's path -> letters -> words -> edit distance
(recipes/slimIPL/src/Train.cpp:236-254, :296-316, :829-872).  Expected values are worked out by hand / by a plain-Python
restatement in this file; the reference tree holds no vectors for these helpers (they live in un-vendored flashlight)."""
import random

import numpy as np
import pytest

LETTERS = "|\n'\n" + "\n".join("abcdefghijklmnopqrstuvwxyz") + "\n"  # WSJ / LibriSpeech letter token set: | ' a..z


def make(criterion="asg", replabel=2, surround="|", lexicon="", usewordpiece=False, tokens=LETTERS):
    from wav2letter_b200.text import TextPipeline

    return TextPipeline(tokens, lexicon, criterion, replabel, surround, usewordpiece, "|")


def idx(ch):
    return {"|": 0, "'": 1}.get(ch, 2 + ord(ch) - ord("a") if ch.isalpha() else None)


def test_dictionary_sizes_follow_the_loop():
    assert make("asg", replabel=2).num_classes == 28 + 2          # + <1> <2>          (Train.cpp:245-247)
    assert make("ctc", replabel=0).num_classes == 28 + 1          # + blank, last      (Train.cpp:248-251)
    assert make("ctc", replabel=1).num_classes == 28 + 1 + 1


def test_target_transform_letters_surround_replabel():
    # conv_glu recipes: --replabel=2 --surround=| ; words fall back to letters with the separator on the right
    tp = make("asg", replabel=2, surround="|")
    got = tp.encode("hello all").tolist()
    R1, R2 = 28, 29
    # h e l l o | a l l |  -> surround: | h e l l o | a l l | | -> replabels: l l -> l <1> ; | | -> | <1>
    want = [idx("|"), idx("h"), idx("e"), idx("l"), R1, idx("o"), idx("|"), idx("a"), idx("l"), R1, idx("|"), R1]
    assert got == want
    # three equal letters -> letter <2>; four -> letter <2> letter
    assert make("asg", 2, "").encode("aaa").tolist() == [idx("a"), R2, idx("|")]
    assert make("asg", 2, "").encode("aaaa").tolist() == [idx("a"), R2, idx("a"), idx("|")]
    # unknown characters are skipped (skipUnk = true at Train.cpp:303)
    assert make("asg", 0, "").encode("a-b").tolist() == [idx("a"), idx("b"), idx("|")]


def test_lexicon_spelling_wins_over_letter_fallback():
    lex = "hello h e l l o |\nhello h e l o |\nworld w o r l d |\n"
    tp = make("ctc", replabel=0, surround="", lexicon=lex)
    assert tp.encode("hello world").tolist() == [idx(c) for c in "hello|world|"]  # first spelling
    assert tp.encode("zz").tolist() == [idx("z"), idx("z"), idx("|")]             # not in the lexicon: letters + separator


def test_batch_padding():
    tp = make("ctc", 0, "")
    t = tp.encode_batch(["ab", "abcd", "a"])
    assert t.shape == (3, 5) and t.dtype == np.int32
    assert t[0].tolist() == [idx("a"), idx("b"), idx("|"), -1, -1] and t[2].tolist() == [idx("a"), idx("|"), -1, -1, -1]


def test_ctc_path_to_letters_and_words():
    tp = make("ctc", replabel=0, surround="")
    blank = tp.num_classes - 1
    path = []
    for ch in "hello|hi|":  # frames: every label twice, blanks between, double letter separated by a blank
        path += [idx(ch), idx(ch), blank]
    ltr = tp.prediction2ltr(path)
    assert ltr == list("hello|hi")  # trailing separator trimmed (tknIdx2Ltr), repeats collapsed, blanks removed
    assert tp.ltr2wrd(ltr) == ["hello", "hi"]
    tgt = tp.encode("hello hi")
    assert tp.target2ltr(np.concatenate([tgt, [-1, -1]])) == list("hello|hi")


def test_asg_path_with_replabels_round_trip():
    tp = make("asg", replabel=2, surround="|")
    for text in ["hello all", "book keeper", "a", "zzz aa"]:
        tgt = tp.encode(text)
        path = np.repeat(tgt, 3)  # a Viterbi path that stays 3 frames on every target position
        assert tp.prediction2ltr(path) == tp.target2ltr(tgt)
        assert "".join(tp.ltr2wrd(tp.prediction2ltr(path))) == text.replace(" ", "")
        assert tp.ltr2wrd(tp.prediction2ltr(path)) == text.split()


def test_wordpiece_tokens_split_into_letters():
    tokens = "_he\nllo\n_wor\nld\n"
    tp = make("ctc", replabel=0, surround="", usewordpiece=True, tokens=tokens)
    from wav2letter_b200.text import TextPipeline

    tp = TextPipeline(tokens, "hello _he llo\nworld _wor ld\n", "ctc", 0, "", True, "_")
    tgt = tp.encode("hello world")
    assert tgt.tolist() == [0, 1, 2, 3]
    blank = tp.num_classes - 1
    ltr = tp.prediction2ltr([0, 0, blank, 1, 2, 2, blank, 3])
    assert ltr == list("hello_world") and tp.ltr2wrd(ltr) == ["hello", "world"]


def lev(a, b):
    d = list(range(len(b) + 1))
    for i in range(1, len(a) + 1):
        prev, d[0] = d[0], i
        for j in range(1, len(b) + 1):
            cur = min(prev + (a[i - 1] != b[j - 1]), d[j] + 1, d[j - 1] + 1)
            prev, d[j] = d[j], cur
    return d[-1]


def test_edit_distance_meter():
    from wav2letter_b200.text import EditDistanceMeter

    m = EditDistanceMeter()
    m.add(list("sitting"), list("kitten"))
    n, ndel, nins, nsub = m.raw()
    assert (n, ndel + nins + nsub) == (6, 3) and nsub == 2 and nins == 1 and ndel == 0
    assert m.value()[0] == pytest.approx(50.0)
    m.add([], list("ab"))       # empty hypothesis: two deletions
    m.add(list("ab"), [])       # empty reference: two insertions
    assert m.raw() == (8, 2, 3, 2)
    rng = random.Random(4)
    for _ in range(200):
        a = [rng.choice("abc") for _ in range(rng.randint(0, 12))]
        b = [rng.choice("abc") for _ in range(rng.randint(0, 12))]
        mm = EditDistanceMeter()
        mm.add(a, b)
        n, ndel, nins, nsub = mm.raw()
        assert n == len(b) and ndel + nins + nsub == lev(a, b) and len(a) - nins + ndel == len(b)


def test_errors_surface():
    from wav2letter_b200 import W2LError
    from wav2letter_b200.text import TextPipeline

    with pytest.raises(W2LError):
        TextPipeline("a\na\n", "", "ctc", 0, "", False, "|")  # duplicate dictionary entry
    with pytest.raises(W2LError):
        make("asg", 0, "", lexicon="").encode_batch([]) if False else TextPipeline(LETTERS, "", "asg", 0, "#", False, "|").encode("a")  # unknown surround
