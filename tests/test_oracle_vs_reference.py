"""CPU, build container only: the oracle against the LIVE, unmodified reference (skipped where /root/reference does not
exist, e.g. on the GPU box).  The committed fixtures are regenerated in memory and compared with what is on disk, so a
stale fixture or a drifting torch version shows up here rather than as a mysterious parity failure."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

pytestmark = pytest.mark.skipif(not os.path.isdir("/root/reference/model"), reason="the reference tree is not present")


def _close(a, b, key):
    if a.dtype.kind in "biu" or b.dtype.kind in "biu":
        assert np.array_equal(a, b), key
    else:
        np.testing.assert_allclose(a, b, atol=1e-6, rtol=1e-5, err_msg=key)


def test_eval_fixture_is_what_the_reference_computes_today(golden_dir):
    import make_golden
    cwd = os.getcwd()
    try:
        flat = make_golden.run_reference("fs2_teacher")
    finally:
        os.chdir(cwd)
    gold = np.load(os.path.join(golden_dir, "fs2_teacher.npz"))
    assert sorted(gold.files) == sorted(flat)
    for k in gold.files:
        _close(np.asarray(flat[k]), gold[k], k)


def test_train_fixture_is_what_the_reference_computes_today(golden_dir):
    import make_golden_train
    cwd = os.getcwd()
    try:
        flat, _ = make_golden_train.run_reference("fs2_train")
    finally:
        os.chdir(cwd)
    gold = np.load(os.path.join(golden_dir, "fs2_train.npz"))
    assert sorted(gold.files) == sorted(flat)
    for k in gold.files:
        _close(np.asarray(flat[k]), gold[k], k)
