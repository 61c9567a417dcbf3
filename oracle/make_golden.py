"""Generate tests/golden/*.npz by running the UNMODIFIED reference on seeded synthetic inputs.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

The reference has no golden vectors of its own for this path, so these
fixtures - outputs of the reference's own classes/functions imported through
``oracle/ref_shim.py`` - are what pins both the oracle and the CUDA kernels.  Conv weights are
snapped to a 2^-12 grid and inputs are spikes / integer counts so every conv partial sum is exact
in fp32 and the fixtures are independent of the conv summation order (bit-exact tier); the
``*_rand`` fixtures keep raw fp32 weights (tolerance tier).
"""
import os

import numpy as np
import torch

from . import ref_shim
from .encodings import synth_events
from .lif import dyadic

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy().copy()


def _save(name, **arrs):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v if isinstance(v, np.ndarray) else np.asarray(v)) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def layer_fixture(ref, name, *, recurrent, Cin, C, B, H, W, T, hard_reset, activation, seed, dyadic_w=True,
                  residual=False, detach=True, int_input=False):
    torch.manual_seed(seed)
    cls = ref.ConvLIFRecurrent if recurrent else ref.ConvLIF
    layer = cls(Cin, C, 3, activation=activation, leak=(0.0, 1.0), thresh=(0.3, 0.1), hard_reset=hard_reset,
                detach=detach)
    with torch.no_grad():
        if dyadic_w:
            layer.ff.weight.copy_(dyadic(layer.ff.weight))
            if recurrent:
                layer.rec.weight.copy_(dyadic(layer.rec.weight))
    g = torch.Generator().manual_seed(seed + 1)
    if int_input:
        xs = torch.poisson(torch.full((T, B, Cin, H, W), 0.3), generator=g)
    else:
        xs = (torch.rand(T, B, Cin, H, W, generator=g) < 0.25).float()
    xs.requires_grad_(True)
    res = (torch.rand(T, B, C, H, W, generator=g) < 0.2).float() if residual else None
    gout = torch.randn(T, B, C, H, W, generator=g)          # d loss / d out[t]
    gv_last = torch.randn(B, C, H, W, generator=g) * 0.1    # d loss / d v[T-1]
    state = None
    vs, zs, outs = [], [], []
    loss = 0
    for t in range(T):
        if recurrent:
            out, state = layer(xs[t], state)
        else:
            out, state = layer(xs[t], state, residual=res[t] if residual else 0)
        vs.append(state[0]); zs.append(state[1]); outs.append(out)
        loss = loss + (out * gout[t]).sum()
    loss = loss + (state[0] * gv_last).sum()
    loss.backward()
    arrs = dict(
        x=_np(xs), w_ff=_np(layer.ff.weight), leak=_np(layer.leak), thresh=_np(layer.thresh),
        lam=_np(torch.sigmoid(layer.leak)), theta=_np(layer.thresh.clamp_min(0.01)),
        v=_np(torch.stack(vs)), z=_np(torch.stack(zs)), out=_np(torch.stack(outs)),
        gout=_np(gout), gv_last=_np(gv_last), g_x=_np(xs.grad), dw_ff=_np(layer.ff.weight.grad),
        dleak=_np(layer.leak.grad), dthresh=_np(layer.thresh.grad),
        meta=np.array([int(recurrent), int(hard_reset), int(detach), int(residual)]), activation=activation,
    )
    if recurrent:
        arrs.update(w_rec=_np(layer.rec.weight), dw_rec=_np(layer.rec.weight.grad))
    if residual:
        arrs.update(residual=_np(res))
    _save(name, **arrs)


def make_net(ref, kind, C, seed, dyadic_w=True, leak=(0.0, 1.0), thresh=(0.3, 0.1)):
    A, AR = ref.adapt(ref.ConvLIF), ref.adapt(ref.ConvLIFRecurrent)
    base = getattr(ref.model, kind)

    class Net(base):
        head_neuron = A
        ff_neuron = A
        rec_neuron = AR if kind == "LIFFireNet" else A

    # model.py never forwards the spiking_neuron dict to the cells, so the
    # active-network leak/thresh statistics are written into the parameters after construction.
    torch.manual_seed(seed)
    cfg = dict(num_bins=2, encoding="cnt", mask_output=False, spiking_neuron=None, base_num_channels=C,
               kernel_size=3, activations=["arctanspike", "arctanspike"], quantization={"enabled": False})
    net = Net(cfg)
    with torch.no_grad():
        for n, p in net.named_parameters():
            if n.endswith(".leak"):
                p.copy_(torch.randn_like(p) * leak[1] + leak[0])
            elif n.endswith(".thresh"):
                p.copy_(torch.randn_like(p) * thresh[1] + thresh[0])
            elif dyadic_w and n.endswith("ff.weight") or n.endswith("rec.weight"):
                p.copy_(dyadic(p))
        net.pred.conv2d.weight.copy_(dyadic(net.pred.conv2d.weight * 20, 10))  # visible, non-saturated flow
    return net


def net_fixture(ref, name, *, kind, C, B, H, W, T, seed):
    net = make_net(ref, kind, C, seed)
    g = torch.Generator().manual_seed(seed + 1)
    cnt = torch.poisson(torch.full((T, B, 2, H, W), 0.25), generator=g)
    flows, acts = [], []
    with torch.no_grad():
        for t in range(T):
            o = net(None, cnt[t].clone(), log=True)
            flows.append(o["flow"][0])
            acts.append([o["activity"][k] for k in sorted(o["activity"])])
    arrs = {"param." + k: _np(v) for k, v in net.state_dict().items()}
    for i, st in enumerate(net._states):
        arrs[f"state{i}"] = _np(st)
    _save(name, cnt=_np(cnt), flow=_np(torch.stack(flows)), activity=np.array(acts, dtype=np.float64),
          kind=kind, **arrs)


def encode_fixture(ref, name, *, n, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    xs, ys, ts, ps = synth_events(n, (H, W), g)
    enc = ref.encodings
    _save(name, xs=_np(xs), ys=_np(ys), ts=_np(ts), ps=_np(ps),
          cnt=_np(enc.events_to_channels(xs, ys, ps, sensor_size=(H, W))),
          voxel5=_np(enc.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(H, W))),
          voxel2=_np(enc.events_to_voxel(xs, ys, ts, ps, 2, sensor_size=(H, W))),
          voxel5_round=_np(enc.events_to_voxel(xs, ys, ts, ps, 5, sensor_size=(H, W), round_ts=True)),
          mask=_np(enc.events_to_image(xs, ys, ps.abs(), sensor_size=(H, W), accumulate=False)),
          img_acc=_np(enc.events_to_image(xs, ys, ps, sensor_size=(H, W), accumulate=True)))


def _event_batch(B, n, H, W, g, T=1):
    ev, pm = [], []
    for _ in range(B):
        xs, ys, ts, ps = synth_events(n, (H, W), g)
        ev.append(torch.stack([ts, ys, xs, ps], dim=1))
        pm.append(torch.stack([(ps > 0).float(), (ps < 0).float()], dim=1))
    return torch.stack(ev), torch.stack(pm)


def iwe_fixture(ref, name, *, B, n, H, W, seed, zero_flow=False, fractional=False):
    g = torch.Generator().manual_seed(seed)
    ev, pm = _event_batch(B, n, H, W, g)
    ev[:, :, 0] *= 3.0                                       # ts in [0,3] like a 3-pass window
    if fractional:
        ev[:, :, 1:3] = ev[:, :, 1:3] * 0.5                  # eval down-scaling path (h5.py:404-405)
    flow = torch.zeros(B, 2, H, W) if zero_flow else torch.tanh(0.5 * torch.randn(B, 2, H, W, generator=g)) * 0.05
    flow.requires_grad_(True)
    iw = ref.iwe
    # per-event flow gather exactly as loss/flow.py:66-81
    fidx = ev[:, :, 1:3].clone()
    fidx[:, :, 0] *= W
    fidx = torch.sum(fidx, dim=2).long()
    fl = flow.view(B, 2, -1)
    ev_flow = torch.cat([torch.gather(fl[:, 1], 1, fidx)[..., None], torch.gather(fl[:, 0], 1, fidx)[..., None]], 2)
    arrs = dict(events=_np(ev), pol_mask=_np(pm), flow=_np(flow), ev_flow=_np(ev_flow),
                params=np.array([H, W, max(H, W)], dtype=np.int64))
    for tag, tref, tsw in (("fw", 3, ev[:, :, 0:1]), ("bw", 0, 3 - ev[:, :, 0:1])):
        idx, w = iw.get_interpolation(ev, ev_flow, tref, (H, W), max(H, W))
        pm4, ts4 = torch.cat([pm] * 4, 1), torch.cat([tsw] * 4, 1)
        imgs = torch.cat([iw.interpolate(idx.long(), w, (H, W), pm4[:, :, 0:1]),
                          iw.interpolate(idx.long(), w, (H, W), pm4[:, :, 1:2]),
                          iw.interpolate(idx.long(), w * ts4, (H, W), pm4[:, :, 0:1]),
                          iw.interpolate(idx.long(), w * ts4, (H, W), pm4[:, :, 1:2])], 1)
        gimg = torch.randn(imgs.shape, generator=g)
        flow.grad = None
        (imgs * gimg).sum().backward(retain_graph=True)
        arrs.update({f"{tag}_idx": _np(idx), f"{tag}_w": _np(w), f"{tag}_img": _np(imgs), f"{tag}_gimg": _np(gimg),
                     f"{tag}_gflow": _np(flow.grad.clone())})
    ev1 = ev.clone()
    ev1[:, :, 0] /= 3.0
    with torch.no_grad():
        arrs["pol_iwe_round"] = _np(iw.compute_pol_iwe(flow, ev1, (H, W), pm[:, :, 0:1], pm[:, :, 1:2], max(H, W), True))
        arrs["pol_iwe_bilinear"] = _np(iw.compute_pol_iwe(flow, ev1, (H, W), pm[:, :, 0:1], pm[:, :, 1:2], max(H, W), False))
    _save(name, **arrs)


def _loader_batch(B, n, H, W, g, enc):
    """One dataloader item in the layout of dataloader/h5.py + base.py:261-278 (custom_collate)."""
    cnt, mask, ev, pm = [], [], [], []
    for _ in range(B):
        xs, ys, ts, ps = synth_events(n, (H, W), g)
        cnt.append(enc.events_to_channels(xs, ys, ps, sensor_size=(H, W)))
        mask.append(enc.events_to_image(xs, ys, ps.abs(), sensor_size=(H, W), accumulate=False)[None])
        ev.append(torch.stack([ts, ys, xs, ps], dim=1))
        pm.append(torch.stack([(ps > 0).float(), (ps < 0).float()], dim=1))
    return torch.stack(cnt), torch.stack(mask), torch.stack(ev), torch.stack(pm)


def train_fixture(ref, name, *, kind, C, B, H, W, T, n, seed, mask_output=False):
    """One optimizer-step worth of the training loop of train_flow.py:232-262 (no clip / Adam)."""
    net = make_net(ref, kind, C, seed)
    net.mask = mask_output
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001}, "model": {"mask_output": mask_output}}
    lossf = ref.flow.EventWarping(cfg, torch.device("cpu"))
    g = torch.Generator().manual_seed(seed + 1)
    arrs = {"param." + k: _np(v) for k, v in net.state_dict().items()}
    flows = []
    for t in range(T):
        cnt, mask, ev, pm = _loader_batch(B, n, H, W, g, ref.encodings)
        arrs.update({f"cnt{t}": _np(cnt), f"mask{t}": _np(mask), f"events{t}": _np(ev), f"pol{t}": _np(pm)})
        out = net(None, cnt)
        out["flow"][0].retain_grad()
        flows.append(out["flow"][0])
        lossf.event_flow_association(out["flow"], ev, pm, mask)
    loss = lossf()
    loss.backward()
    arrs["loss"] = _np(loss)
    arrs["flow"] = _np(torch.stack(flows))
    arrs["gflow"] = _np(torch.stack([f.grad for f in flows]))
    for k, p in net.named_parameters():
        arrs["grad." + k] = _np(p.grad)
    _save(name, kind=kind, dims=np.array([C, B, H, W, T, n]), mask_output=int(mask_output), **arrs)



def train_step_fixture(ref, name, *, C, B, H, W, T, n, seed, default_params=False, lr=2e-4, clip=1.0):
    """One FULL optimizer step at a benchmark shape (train_flow.py:232-279: T bins, EventWarping, backward,
    clip_grad_norm_, Adam): loss, every parameter gradient (before clipping), the total gradient norm and the updated
    parameters.  The window is regenerated from its seed by tests/snnflow_testutil.synth_window (a checksum guards the
    generator); `default_params`: the cells' own init statistics and raw fp32 weights (a silent last layer: flow == 0,
    every event on the integer-tie path of utils/iwe.py:57-59), else the active dyadic network of make_net."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(OUT)))
    from snnflow_testutil import synth_window
    torch.set_num_threads(max(torch.get_num_threads(), os.cpu_count() or 1))
    if default_params:
        net = make_net(ref, "LIFFireNet", C, seed, dyadic_w=False, leak=(-4.0, 0.1), thresh=(0.8, 0.0))
        with torch.no_grad():
            net.pred.conv2d.weight.mul_(1.0 / 20)   # back to ~w_scale_pred (model.py:43)
    else:
        net = make_net(ref, "LIFFireNet", C, seed)
    w = synth_window(T, B, n, H, W, seed + 1)
    arrs = {"param." + k: _np(v) for k, v in net.state_dict().items()}
    arrs["lam"] = _np(torch.stack([torch.sigmoid(getattr(net, l).leak).reshape(-1) for l in
                                   ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b")]))
    arrs["theta"] = _np(torch.stack([getattr(net, l).thresh.clamp_min(0.01).reshape(-1) for l in
                                     ("head", "G1", "R1a", "R1b", "G2", "R2a", "R2b")]))
    cfg = {"loader": {"resolution": [H, W]}, "loss": {"flow_regul_weight": 0.001}, "model": {"mask_output": False}}
    lossf = ref.flow.EventWarping(cfg, torch.device("cpu"))
    opt = torch.optim.Adam(net.parameters(), lr=lr)
    flows = []
    for t in range(T):
        out = net(None, w["event_cnt"][t])
        out["flow"][0].retain_grad()
        flows.append(out["flow"][0])
        lossf.event_flow_association(out["flow"], w["event_list"][t].clone(), w["event_list_pol_mask"][t], w["event_mask"][t])
    loss = lossf()
    loss.backward(retain_graph=True)
    for k, p in net.named_parameters():
        arrs["grad." + k] = _np(p.grad)
    gflow = torch.stack([f.grad for f in flows])
    names = [k for k, _ in net.named_parameters()]
    plist = [p for _, p in net.named_parameters()]
    # (a) BPTT alone, on a well-conditioned loss: sum(flow * G), G ~ N(0, 1) regenerated from the seed by the test
    G = torch.randn(T, B, 2, H, W, generator=torch.Generator().manual_seed(seed + 2))
    lin = torch.autograd.grad([f for f in flows], plist, grad_outputs=[G[t] for t in range(T)], retain_graph=True)
    for k, gk in zip(names, lin):
        arrs["gradlin." + k] = _np(gk)
    # (b) how well is d loss / d flow defined in fp32?  The same loss arithmetic in float64 (oracle/loss.py restates
    # loss/flow.py and reproduces it bit for bit in fp32) on the SAME fp32 flow maps gives the exact gradient; pixels
    # holding ~1e-6 of bilinear weight divide by (count + 1e-9) (loss/flow.py:214-217) and lose every digit in fp32.
    from .loss import EventWarpingOracle
    f64 = torch.stack([f.detach() for f in flows]).double().requires_grad_(True)
    l64 = EventWarpingOracle((H, W), 0.001)
    for t in range(T):
        l64.associate(f64[t], w["event_list"][t].clone().double(), w["event_list_pol_mask"][t].double(), w["event_mask"][t].double())
    loss64 = l64()
    loss64.backward()
    g64 = f64.grad
    arrs["gflow_relerr_fp32"] = np.array([float((gflow[t].double() - g64[t]).norm() / g64[t].norm()) for t in range(T)])
    # ... and the reference's own (fp32) BPTT applied to that exact loss gradient: the parameter gradients an exact loss
    # backward would have produced
    ex = torch.autograd.grad([f for f in flows], plist, grad_outputs=[g64[t].float() for t in range(T)])
    for k, gk in zip(names, ex):
        arrs["grad64." + k] = _np(gk)
    arrs["loss64"] = np.array(float(loss64))
    total = torch.nn.utils.clip_grad.clip_grad_norm_(net.parameters(), clip)
    opt.step()
    for k, p in net.named_parameters():
        arrs["new." + k] = _np(p)
    arrs.update(loss=_np(loss), grad_norm=_np(total), flow_last=_np(flows[-1]), gflow_last=_np(gflow[-1]),
                flow_absmax=np.array([float(f.detach().abs().max()) for f in flows]),
                gflow_norm=np.array([float(g.norm()) for g in gflow]),
                spike_rate=np.array([float(st[1].mean()) for st in net._states]),
                window_checksum=np.array([float(w["event_cnt"].double().sum()), float(w["event_list"].double().sum()),
                                          float((w["event_list"][..., 0].double() * w["event_list"][..., 2].double()).sum())]))
    _save(name, dims=np.array([C, B, H, W, T, n]), seed=seed, lr=lr, clip=clip, default_params=int(default_params), **arrs)


# ------------------------------------------------------------------------------------------------
# loader fixtures: the reference's own H5Loader.__getitem__ driven through an in-memory stand-in for h5py
# ------------------------------------------------------------------------------------------------
class _FakeDataset:
    """h5py.Dataset stand-in: slicing returns a COPY (the loader subtracts t0 in place, h5.py:129)."""

    def __init__(self, arr, attrs=None):
        self.arr, self.attrs, self.dtype = arr, attrs or {}, arr.dtype

    def __getitem__(self, k):
        return np.array(self.arr[k], copy=True)

    def __len__(self):
        return len(self.arr)


class _FakeGroup:
    def __init__(self, items):
        self.items = items

    def visititems(self, cb):
        for k, v in self.items.items():
            cb(k, v)

    def __getitem__(self, k):
        return self.items[k]


class _FakeFile:
    registry = {}

    def __init__(self, path, mode="r"):
        self.d = _FakeFile.registry[os.path.basename(path)]
        self.attrs = self.d["attrs"]

    def __getitem__(self, k):
        return self.d[k]

    def close(self):
        pass


def _raw_stream(n, H, W, g, hot_px=()):
    """A raw sensor stream as stored in the HDF5 files: integer coords, seconds, polarity in {0,1}; a few pixels fire in
    (almost) every window so the hot-pixel filter has something to remove."""
    xs = torch.randint(0, W, (n,), generator=g).numpy().astype(np.int16)
    ys = torch.randint(0, H, (n,), generator=g).numpy().astype(np.int16)
    for j, (hy, hx) in enumerate(hot_px):
        xs[j::37], ys[j::37] = hx, hy
    ts = 10.0 + np.sort(torch.rand(n, generator=g).double().numpy()) * 0.5
    ps = torch.randint(0, 2, (n,), generator=g).numpy().astype(np.int8)
    return xs, ys, ts, ps


def loader_fixture(ref, name, *, mode, B, H, W, n_win, n_items, num_bins, round_enc, seed, augment_prob=(0.5, 0.5, 0.5),
                   hot=None, target=None):
    import importlib
    import tempfile
    import types
    h5 = importlib.import_module("dataloader.h5")
    h5.h5py = types.SimpleNamespace(File=_FakeFile)
    g = torch.Generator().manual_seed(seed)
    tmp = tempfile.mkdtemp()
    arrs = {}
    _FakeFile.registry = {}
    for b in range(B):
        fn = f"seq{b}.h5"
        open(os.path.join(tmp, fn), "w").close()
        n_tot = n_win * (n_items + 2)
        xs, ys, ts, ps = _raw_stream(n_tot, H, W, g, hot_px=[(1, 2), (H - 2, W - 3), (3, 3)] if hot else ())
        d = {"events/xs": _FakeDataset(xs), "events/ys": _FakeDataset(ys), "events/ts": _FakeDataset(ts),
             "events/ps": _FakeDataset(ps), "attrs": {"t0": 10.0, "duration": 0.5}}
        if mode == "gtflow_dt1":
            # ground-truth maps every n_win events: the loader windows the stream between their timestamps
            maps = {}
            for i in range(n_items + 2):
                fm = torch.randn(2, H, W, generator=g).numpy().astype(np.float32)
                maps[f"{i:06d}"] = _FakeDataset(fm, {"timestamp": float(ts[min(i * n_win, n_tot - 1)])})
            d["flow_dt1"] = _FakeGroup(maps)
            for k, v in maps.items():
                d["flow_dt1/" + k] = v
            d["flow_dt1"].items = maps
        _FakeFile.registry[fn] = d
        d["raw"] = (xs, ys, ts, ps)
    cfg = {"data": {"mode": mode, "window": n_win if mode == "events" else 1, "path": tmp},
           "loader": {"resolution": list(target or (H, W)), "std_resolution": [H, W], "batch_size": B,
                      "augment": ["Horizontal", "Vertical", "Polarity"], "augment_prob": list(augment_prob)},
           "hot_filter": dict(enabled=bool(hot), **(hot or dict(max_px=100, min_obvs=5, max_rate=0.8))),
           "vis": {"bars": False}}
    if mode == "events":
        cfg["loader"]["resolution"] = [H, W]
    np.random.seed(seed)
    loader = h5.H5Loader(cfg, num_bins, round_encoding=round_enc)
    for b in range(B):   # os.walk order decides which file feeds which batch slot (h5.py:59-69)
        raw = _FakeFile.registry[os.path.basename(loader.files[b])]["raw"]
        arrs.update({f"raw{b}.{k}": v for k, v in zip(("xs", "ys", "ts", "ps"), raw)})
    arrs["flips"] = np.array([[loader.batch_augmentation[m][b] for m in ("Horizontal", "Vertical", "Polarity")]
                              for b in range(B)], dtype=np.int32)
    for it in range(n_items):
        items = [loader[b] for b in range(B)]
        batch = loader.custom_collate(items)
        if mode != "events":
            # the window the loader cut out of the stream (indices into the raw arrays), to replay it elsewhere
            f_obj = loader.open_files[0]
            i0 = loader.find_ts_index(f_obj, loader.open_files_flowmaps[0].ts[it])
            i1 = loader.find_ts_index(f_obj, loader.open_files_flowmaps[0].ts[it + 1])
            arrs[f"item{it}.range"] = np.array([i0, i1])
        for k in ("event_cnt", "event_voxel", "event_mask", "event_list", "event_list_pol_mask"):
            arrs[f"item{it}.{k}"] = _np(batch[k])
    dims = dict(B=B, H=H, W=W, n_win=n_win, n_items=n_items, num_bins=num_bins, round_enc=int(round_enc))
    _save(name, mode=mode, hot=np.array([hot["max_px"], hot["min_obvs"], hot["max_rate"]]) if hot else np.zeros(0),
          target=np.array(target or (H, W)), **{k: np.array(v) for k, v in dims.items()}, **arrs)

def main(only=None):
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    ref = ref_shim.load()
    if only:   # regenerate selected fixtures only (python oracle/make_golden.py --only name[,name])
        g = globals()
        for fn in ("layer_fixture", "net_fixture", "encode_fixture", "iwe_fixture", "train_fixture", "train_step_fixture",
                   "loader_fixture"):
            orig = g[fn]
            g[fn] = (lambda o: (lambda r, name, **kw: o(r, name, **kw) if name in only else None))(orig)
    # --- single layers (fwd + bwd), bit-exact tier (dyadic weights, spike inputs) ---
    layer_fixture(ref, "layer_ff_hard_arctan", recurrent=False, Cin=4, C=8, B=2, H=11, W=13, T=4,
                  hard_reset=True, activation="arctanspike", seed=10)
    layer_fixture(ref, "layer_ff_soft_super_res", recurrent=False, Cin=3, C=6, B=2, H=9, W=35, T=4,
                  hard_reset=False, activation="superspike", seed=11, residual=True)
    layer_fixture(ref, "layer_rec_hard_arctan", recurrent=True, Cin=8, C=8, B=2, H=12, W=10, T=4,
                  hard_reset=True, activation="arctanspike", seed=12)
    layer_fixture(ref, "layer_rec_soft_triangle", recurrent=True, Cin=5, C=8, B=1, H=8, W=8, T=4,
                  hard_reset=False, activation="trianglespike", seed=13)
    layer_fixture(ref, "layer_rec_hard_mgspike", recurrent=True, Cin=6, C=8, B=2, H=10, W=12, T=4,
                  hard_reset=True, activation="mgspike", seed=19)
    layer_fixture(ref, "layer_rec_nodetach", recurrent=True, Cin=4, C=4, B=1, H=8, W=9, T=3,
                  hard_reset=True, activation="arctanspike", seed=14, detach=False)
    layer_fixture(ref, "layer_head_counts", recurrent=False, Cin=2, C=32, B=2, H=16, W=20, T=3,
                  hard_reset=True, activation="arctanspike", seed=15, int_input=True)
    layer_fixture(ref, "layer_ff_c32", recurrent=False, Cin=32, C=32, B=1, H=18, W=40, T=3,
                  hard_reset=True, activation="arctanspike", seed=16)
    layer_fixture(ref, "layer_rec_c32", recurrent=True, Cin=32, C=32, B=2, H=16, W=24, T=3,
                  hard_reset=True, activation="arctanspike", seed=17)
    layer_fixture(ref, "layer_rec_c32_rand", recurrent=True, Cin=32, C=32, B=1, H=16, W=16, T=2,
                  hard_reset=True, activation="arctanspike", seed=18, dyadic_w=False)
    # --- whole networks, T bins (bit-exact spikes / membranes) ---
    net_fixture(ref, "net_firenet_c8", kind="LIFFireNet", C=8, B=2, H=16, W=16, T=6, seed=20)
    net_fixture(ref, "net_fireflownet_c8", kind="LIFFireFlowNet", C=8, B=2, H=16, W=24, T=4, seed=21)
    net_fixture(ref, "net_firenet_c32", kind="LIFFireNet", C=32, B=1, H=16, W=16, T=5, seed=22)
    # --- encodings ---
    encode_fixture(ref, "encode_small", n=4000, H=24, W=32, seed=30)
    encode_fixture(ref, "encode_empty", n=0, H=8, W=8, seed=31)
    # --- IWE ---
    iwe_fixture(ref, "iwe_rand", B=2, n=600, H=16, W=20, seed=40)
    iwe_fixture(ref, "iwe_zero_flow", B=1, n=200, H=12, W=12, seed=41, zero_flow=True)
    iwe_fixture(ref, "iwe_fractional", B=1, n=300, H=16, W=16, seed=42, fractional=True)
    # --- loss + one training window ---
    train_fixture(ref, "train_firenet_c8", kind="LIFFireNet", C=8, B=2, H=16, W=16, T=3, n=120, seed=50)
    train_fixture(ref, "train_fireflownet_c8_mask", kind="LIFFireFlowNet", C=8, B=1, H=16, W=16, T=3, n=150,
                  seed=51, mask_output=True)
    # C = 16 / 32: inside the envelope of the layer-major window engine (tensor-core head layer included)
    train_fixture(ref, "train_firenet_c16", kind="LIFFireNet", C=16, B=2, H=16, W=16, T=3, n=120, seed=52)
    train_fixture(ref, "train_fireflownet_c32", kind="LIFFireFlowNet", C=32, B=1, H=12, W=20, T=4, n=150, seed=53)
    # --- one full optimizer step at BASELINE.json configs[1] (C=32, batch 8, 128x128, 10 bins x 1000 events) ---
    train_step_fixture(ref, "step_cfg1_active", C=32, B=8, H=128, W=128, T=10, n=1000, seed=70)
    train_step_fixture(ref, "step_cfg1_default", C=32, B=8, H=128, W=128, T=10, n=1000, seed=71, default_params=True)
    # --- loader: raw event windows -> batch tensors (the reference's H5Loader on an in-memory stream) ---
    loader_fixture(ref, "loader_events_hot", mode="events", B=3, H=20, W=24, n_win=400, n_items=5, num_bins=5,
                   round_enc=False, seed=60, hot=dict(max_px=2, min_obvs=2, max_rate=0.7))
    loader_fixture(ref, "loader_events_round", mode="events", B=2, H=16, W=16, n_win=300, n_items=2, num_bins=2,
                   round_enc=True, seed=61, augment_prob=(1.0, 0.0, 1.0))
    loader_fixture(ref, "loader_gtflow_pool", mode="gtflow_dt1", B=1, H=32, W=48, n_win=500, n_items=3, num_bins=3,
                   round_enc=False, seed=62, target=(16, 24), hot=dict(max_px=100, min_obvs=1, max_rate=0.9))


if __name__ == "__main__":
    import sys
    sel = None
    if "--only" in sys.argv:
        sel = set(sys.argv[sys.argv.index("--only") + 1].split(","))
    main(sel)
