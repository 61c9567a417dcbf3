"""Oracle (TEST INFRASTRUCTURE): the loader's event branch restated a second time, in plain numpy (no torch ops), for the
outputs that are exact by construction - per-polarity counts, the any-event mask, the event list and the polarity mask -
including flips, the hot-pixel filter and the average-pool down-sampling.  Follows dataloader/base.py:71-126,160-256,
dataloader/encodings.py:30-45,70-103 and dataloader/h5.py:323-331,375-410 like ``oracle/loader.py`` does, but shares no
code with it: the two restatements are checked against each other and against the fixtures written by the reference's
own ``H5Loader`` (tests/test_oracle_vs_golden.py)."""
import numpy as np

F32 = np.float32


def format_item_np(xs, ys, ts, ps, *, resolution, flips=(False, False, False), hot_state=None, hot_cfg=None, target=None):
    """xs, ys: integer sensor coordinates; ts: float32 seconds relative to t0; ps: raw polarity in {0,1}.
    hot_state: dict(events=[H,W] float32 array, idx=int) updated in place when hot_cfg is given.
    Returns dict(event_cnt [2,h,w], event_mask [1,h,w], event_list [4,N], event_list_pol_mask [2,N])."""
    H, W = resolution
    xs, ys = xs.astype(F32), ys.astype(F32)
    ts = ts.astype(F32)
    p = ps.astype(F32) * F32(2) - F32(1)                                   # base.py:89
    if ts.size:                                                             # base.py:90-98
        rng = F32(ts.max() - ts.min())
        ts = ((ts - ts.min()) / rng).astype(F32) if rng > 0 else np.zeros_like(ts)
    if flips[0]:
        xs = F32(W - 1) - xs                                               # base.py:114-116
    if flips[1]:
        ys = F32(H - 1) - ys                                               # base.py:118-120
    if flips[2]:
        p = p * F32(-1)                                                    # base.py:122-124
    xi, yi = xs.astype(np.int64), ys.astype(np.int64)
    cnt = np.zeros((2, H, W), dtype=np.int64)                              # encodings.py:70-85: +1 per event on its polarity plane
    np.add.at(cnt[0], (yi[p > 0], xi[p > 0]), 1)
    np.add.at(cnt[1], (yi[p < 0], xi[p < 0]), 1)
    cnt = cnt.astype(F32)
    mask = np.zeros((H, W), dtype=F32)                                     # encodings.py:43, accumulate=False: the last event wins
    for k in range(len(xi)):
        mask[yi[k], xi[k]] = abs(p[k])
    ev_list = np.stack([ts, ys, xs, p]).astype(F32)                        # base.py:221
    pol = np.stack([np.where(p < 0, F32(0), p), np.where(p > 0, F32(0), p) * F32(-1)]).astype(F32)   # base.py:231-235
    if hot_cfg is not None:                                                # base.py:246-256, encodings.py:88-103
        hot_state["events"] += ((cnt[0] + cnt[1]) > 0).astype(F32)
        hot_state["idx"] += 1
        idx = hot_state["idx"]
        hm = np.ones((H, W), dtype=F32)
        if idx > hot_cfg["min_obvs"]:
            rate = (hot_state["events"] / F32(idx)).astype(F32)
            order = sorted(range(H * W), key=lambda i: (-rate.flat[i], i))   # argmax after argmax: rate descending, index ascending
            for i in order[:hot_cfg["max_px"]]:
                if rate.flat[i] > F32(hot_cfg["max_rate"]):
                    hm.flat[i] = 0
                else:
                    break
        cnt = cnt * hm
        mask = mask * hm
    mask = mask[None]
    if target is not None and (target[0] < H or target[1] < W):           # h5.py:375-410
        th, tw = target
        kh, kw = H // th, W // tw
        oh, ow = H // kh, W // kw

        def pool(img):
            c = img.shape[0]
            v = img[:, :oh * kh, :ow * kw].reshape(c, oh, kh, ow, kw).transpose(0, 1, 3, 2, 4).reshape(c, oh, ow, kh * kw)
            acc = np.zeros((c, oh, ow), dtype=F32)
            for j in range(kh * kw):                                       # row-major window sum in fp32, one division
                acc = (acc + v[..., j]).astype(F32)
            return (acc / F32(kh * kw)).astype(F32)

        cnt, mask = pool(cnt), pool(mask)
        if ev_list.size:
            ev_list = ev_list.copy()
            ev_list[1] = np.clip((ev_list[1] * F32(th / H)).astype(F32), 0, th - 1)
            ev_list[2] = np.clip((ev_list[2] * F32(tw / W)).astype(F32), 0, tw - 1)
    return {"event_cnt": cnt, "event_mask": mask, "event_list": ev_list, "event_list_pol_mask": pol}
