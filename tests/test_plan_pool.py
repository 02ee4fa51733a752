"""CPU test of the shim's plan-buffer handoff (fbtt_embedding_b200/tt_embeddings.py:_plan_for): host
bookkeeping only, CPU tensors stand in for device buffers, no kernel is launched."""
import pytest
import torch


@pytest.fixture()
def ext(monkeypatch):
    from fbtt_embedding_b200 import tt_embeddings as e

    capturing = {"on": False}
    monkeypatch.setattr(torch.cuda, "is_current_stream_capturing", lambda: capturing["on"])
    e._plan_cache.clear()
    e._plan_free.clear()
    e._capturing = capturing
    yield e
    e._plan_cache.clear()
    e._plan_free.clear()
    del e._capturing


def _batch(n=100):
    return [torch.zeros(n, dtype=torch.int64) for _ in range(3)]


def test_forward_parks_backward_retires_next_forward_reuses(ext):
    sh = ext._shape(1, 8, 64, [200, 220, 250], [4, 4, 4], [1, 32, 32, 1])
    nb = ext._workspace_bytes(sh, 100)
    assert nb > 0
    i, r, t = _batch()
    plan, ready, key = ext._plan_for(sh, 100, i, r, t, nb, True, 7)
    assert ready == 0 and len(ext._plan_cache) == 1
    hb = ext._lib.ttb_tt_workspace_header_bytes(__import__("ctypes").byref(sh), 100)
    assert int(plan[:hb].abs().sum()) == 0  # header contract: zero on entry
    again, ready2, key2 = ext._plan_for(sh, 100, i, r, t, nb, False, 7)
    assert ready2 == 1 and again is plan and key2 == key
    ext._plan_done(key2, True)
    assert not ext._plan_cache and ext._plan_free[(None, 7, hb)] == [plan]
    # fresh tensors, same stream: pooled buffer, no allocation; another stream does not see it
    i2, r2, t2 = _batch()
    other, _, key_o = ext._plan_for(sh, 100, i2, r2, t2, nb, True, 9)
    assert other is not plan
    reused, ready3, key3 = ext._plan_for(sh, 100, i2, r2, t2, nb, True, 7)
    assert reused is plan and ready3 == 0 and not ext._plan_free[(None, 7, hb)]
    # an in-place edit of the indices invalidates the plan (version counter in the key)
    i2.add_(1)
    _, ready4, key4 = ext._plan_for(sh, 100, i2, r2, t2, nb, False, 7)
    assert ready4 == 0 and key4 != key3


def test_failed_call_drops_buffer_and_capture_buffers_stay_out_of_the_pool(ext):
    sh = ext._shape(1, 8, 64, [200, 220, 250], [4, 4, 4], [1, 32, 32, 1])
    nb = ext._workspace_bytes(sh, 100)
    hb = ext._lib.ttb_tt_workspace_header_bytes(__import__("ctypes").byref(sh), 100)
    i, r, t = _batch()
    _, _, key = ext._plan_for(sh, 100, i, r, t, nb, True, 0)
    ext._plan_done(key, False)
    assert not ext._plan_cache and not any(ext._plan_free.values())
    ext._capturing["on"] = True
    _, _, key = ext._plan_for(sh, 100, i, r, t, nb, True, 0)
    assert ext._plan_cache[key][4] is False
    ext._plan_done(key, True)
    assert not any(ext._plan_free.values())
    # while capturing, pooled eager buffers are not baked into the graph either
    ext._capturing["on"] = False
    plan, _, key = ext._plan_for(sh, 100, i, r, t, nb, True, 0)
    ext._plan_done(key, True)
    ext._capturing["on"] = True
    fresh, _, _ = ext._plan_for(sh, 100, i, r, t, nb, True, 0)
    assert fresh is not plan and ext._plan_free[(None, 0, hb)] == [plan]


def test_aged_out_entries_are_retired_to_the_pool(ext):
    sh = ext._shape(1, 8, 64, [200, 220, 250], [4, 4, 4], [1, 32, 32, 1])
    nb = ext._workspace_bytes(sh, 100)
    keep = []
    for _ in range(70):  # inference: forwards only, nothing ever consumes the plans
        b = _batch()
        keep.append(b)
        ext._plan_for(sh, 100, *b, nb, True, 0)
    assert len(ext._plan_cache) == 64
    assert 0 < sum(len(v) for v in ext._plan_free.values()) <= 8


def test_a_plan_with_a_larger_header_never_gets_a_buffer_with_a_smaller_clean_prefix(ext):
    small = ext._shape(1, 8, 64, [200, 220, 250], [4, 4, 4], [1, 32, 32, 1])
    big = ext._shape(3, 8, 64, [200, 220, 250], [4, 4, 4], [1, 32, 32, 1])  # 3 tables: 3x the bucket counters
    i, r, t = _batch()
    plan, _, key = ext._plan_for(small, 100, i, r, t, ext._workspace_bytes(small, 100), True, 0)
    plan.fill_(0xAB)  # what the body looks like after use; the kernels would have re-zeroed the small header only
    ext._plan_done(key, True)
    other, _, _ = ext._plan_for(big, 100, i, r, t, ext._workspace_bytes(big, 100), True, 0)
    hb = ext._lib.ttb_tt_workspace_header_bytes(__import__("ctypes").byref(big), 100)
    assert other is not plan and int(other[:hb].abs().sum()) == 0
