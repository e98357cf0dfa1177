"""Host-side logic of the densification bookkeeping that needs no GPU: the schedule rule of `adaptive_control`
(/root/reference/networks/gaussian_splatting.py:667-703 -> my_ext/utils/utils.py:126-146 `check_interval_v2`)."""
import importlib.util
import os

import pytest

from sk_gs_b200.densify import DEFAULT_CONTROL, check_interval

REF_UTILS = '/root/reference/my_ext/utils/utils.py'


def test_default_schedule():
    """exps/default.yaml:65-74: densify / prune every 100 steps inside (500, 25000), opacity reset every 3000 after 3000."""
    di = DEFAULT_CONTROL['densify_interval']
    hits = [s for s in range(0, 26000) if check_interval(s, *di)]
    assert hits[0] == 600 and hits[-1] == 24_900 and len(hits) == 244
    oi = DEFAULT_CONTROL['opacity_reset_interval']
    assert [s for s in range(0, 10000) if check_interval(s, *oi)] == [6000, 9000]
    assert not check_interval(100, 0, 0, -1) and not check_interval(100, -5, 0, -1)


@pytest.mark.skipif(not os.path.exists(REF_UTILS), reason='the reference tree is only present in the authoring container')
def test_matches_reference_check_interval_v2():
    src = open(REF_UTILS).read()
    start = src.index('def check_interval_v2')
    end = src.index('\ndef ', start + 10)
    ns = {}
    exec(src[start:end], ns)  # the function is self-contained (no imports)
    ref = ns['check_interval_v2']
    for interval, a, b in ((100, 500, 25_000), (3000, 3000, -1), (100, 0, -1), (7, None, None), (0, 0, 10), (5, -1, 23)):
        for step in range(0, 200 if interval < 50 else 26_000, 1 if interval < 50 else 50):
            assert check_interval(step, interval, a, b) == ref(step, interval, a, b, close='()'), (step, interval, a, b)
