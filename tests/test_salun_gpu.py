"""GPU parity tests of the SalUn saliency mask and the masked AdamW step (through the C ABI).
Integer / byte results (accumulated bf16 bit patterns, thresholds, int8 masks) must be bit-exact."""
import numpy as np
import pytest
import torch

from oracle import salun_oracle as SO
from tests._golden import GOLDEN

pytestmark = pytest.mark.gpu


def _flat(layout, named: dict, dtype):
    buf = torch.zeros(layout.total, dtype=dtype)
    for i, n in enumerate(layout.names):
        buf[layout.offsets[i]: layout.offsets[i] + layout.numels[i]] = named[n].reshape(-1)
    return buf


def _check_masks(state, named_acc, thresholds):
    want = SO.masks(named_acc, thresholds)
    got = dict(state.masks(thresholds))
    for th in thresholds:
        for n in named_acc:
            g = got[th][n].cpu().numpy()
            assert g.dtype == np.int8 and g.shape == tuple(named_acc[n].shape)
            assert np.array_equal(g, want[th][n]), (th, n, int((g != want[th][n]).sum()))


def test_reference_fixture_bit_exact():
    """Per-batch gradients recorded from the reference's save_gradient_ratio -> same ten masks."""
    from multimodal_unlearning_b200 import salun
    z = np.load(GOLDEN / "salun_tiny.npz")
    names = [str(n) for n in z["names"]]
    shapes = [tuple(int(x) for x in str(s).split(";")) for s in z["shapes"]]
    layout = salun.FlatLayout(list(zip(names, shapes)))
    offs = np.concatenate([[0], np.cumsum(layout.numels)])
    state = salun.SaliencyState(layout, "cuda:0")
    acc = None
    for b in range(3):
        g = torch.from_numpy(z[f"grad_{b}"].copy()).view(torch.bfloat16)
        named = {n: g[offs[i]:offs[i + 1]] for i, n in enumerate(names)}
        state.accumulate_flat(_flat(layout, named, torch.bfloat16).cuda())
        acc = g.clone() if acc is None else SO.accumulate(acc, g)
    named_acc = {n: acc[offs[i]:offs[i + 1]].reshape(shapes[i]) for i, n in enumerate(names)}
    got_acc = state.acc.cpu()
    for i, n in enumerate(names):
        a = got_acc[layout.offsets[i]: layout.offsets[i] + layout.numels[i]]
        assert torch.equal(a.view(torch.int16), named_acc[n].reshape(-1).view(torch.int16)), n
    n_total = int(offs[-1])
    for th, hard in state.masks(SO.REFERENCE_THRESHOLDS):
        flat = np.concatenate([hard[n].reshape(-1).cpu().numpy() for n in names])
        stable = np.unpackbits(z[f"mask_stable_{th}"])[:n_total].astype(np.int8)
        assert np.array_equal(flat, stable), th


def test_ties_odd_sizes_and_special_values():
    from multimodal_unlearning_b200 import salun
    g = torch.Generator().manual_seed(0)
    shapes = {"a": (1000003,), "b": (8,), "c": (3, 5), "d": (2048, 1024), "e": (1,), "f": (70000,)}
    named = {}
    for n, s in shapes.items():
        x = torch.randn(*s, generator=g)
        named[n] = x
    named["a"] = (named["a"] * 4).round() / 4            # heavy ties
    named["d"][:, ::3] = 0.0                              # a third exact zeros
    named["d"][5, 7] = 1e6                                # overflows fp16 -> inf
    named["d"][6, 7] = -7e4
    named["f"] = named["f"] * 1e-9                        # flushes to fp16 zero / subnormals
    named = {n: v.to(torch.bfloat16) for n, v in named.items()}
    layout = salun.FlatLayout([(n, v.shape) for n, v in named.items()])
    state = salun.SaliencyState(layout, "cuda:0")
    state.accumulate_flat(_flat(layout, named, torch.bfloat16).cuda())
    ths = (0.0, 0.1, 0.25, 0.5, 0.9, 0.999, 1.0)
    _check_masks(state, named, ths)
    sel = state.select(ths).cpu().numpy()
    for i, n in enumerate(layout.names):
        keys = SO.saliency_keys(named[n])
        pat = torch.arange(0x8000, dtype=torch.int16).view(torch.bfloat16)
        key_of_pattern = SO.saliency_keys(pat)        # key of every bf16 magnitude bit pattern
        for j, th in enumerate(ths):
            tau, ties = SO.select(keys, th)
            assert (int(sel[i, j, 0]), int(sel[i, j, 1])) == (tau, ties), (n, th)
            lo, hi = int(sel[i, j, 2]), int(sel[i, j, 3])
            if tau < 0:
                assert lo > 0x7FFF and hi >= 0x7FFF   # nothing above hi, nothing inside [lo, hi]
            else:
                inside = np.nonzero(key_of_pattern == tau)[0]
                assert (lo, hi) == (int(inside.min()), int(inside.max())), (n, th)


def test_accumulate_per_tensor_matches_flat():
    from multimodal_unlearning_b200 import salun
    model = torch.nn.Sequential(torch.nn.Linear(64, 48), torch.nn.Linear(48, 10)).to(torch.bfloat16).cuda()
    layout = salun.FlatLayout.of_model(model)
    state = salun.SaliencyState(layout, "cuda:0")
    acc = {n: None for n in layout.names}
    g = torch.Generator().manual_seed(1)
    for _ in range(3):
        for n, p in model.named_parameters():
            p.grad = torch.randn(p.shape, generator=g).to(torch.bfloat16).cuda()
            gr = p.grad.cpu()
            acc[n] = gr.clone() if acc[n] is None else SO.accumulate(acc[n], gr)
        state.accumulate_model(model)
    for i, n in enumerate(layout.names):
        assert torch.equal(layout.view(state.acc, i).cpu().view(torch.int16), acc[n].view(torch.int16))
    _check_masks(state, acc, (0.3, 0.5))


def test_large_properties():
    """Sizes far beyond what the CPU sort finishes in seconds: order-statistic properties instead."""
    from multimodal_unlearning_b200 import salun
    dev = torch.device("cuda:0")
    shapes = [("w0", (4096, 11008)), ("w1", (32000, 4096)), ("n0", (4096,)), ("w2", (11008, 4096))]
    layout = salun.FlatLayout(shapes)
    state = salun.SaliencyState(layout, dev)
    g = torch.Generator(device=dev).manual_seed(2)
    grads = (1e-3 * torch.randn(layout.total, generator=g, device=dev)).to(torch.bfloat16)
    state.accumulate_flat(grads)
    assert torch.equal(state.acc, grads)
    ths = (0.1, 0.5, 0.9, 1.0)
    sel = state.select(ths)
    prev = None
    keys_all = state.acc.float().abs().to(torch.float16).view(torch.int16).to(torch.int32)
    for j, th in enumerate(ths):
        flat = state.mask_flat(sel, j)
        for i, (n, s) in enumerate(shapes):
            m = layout.view(flat, i).reshape(-1)
            k = layout.view(keys_all, i).reshape(-1)
            cnt = int(m.sum().item())
            assert cnt == int(th * m.numel()), (n, th)                      # exactly k ones
            if 0 < cnt < m.numel():
                lo_sel = k[m == 1].min().item()
                hi_unsel = k[m == 0].max().item()
                assert lo_sel >= hi_unsel                                   # a top-k set
                tau = int(sel[i, j, 0].item())
                assert lo_sel == tau
                # ties: selected tau-elements come before unselected ones (stable order)
                tie = (k == tau)
                pos = torch.nonzero(tie).reshape(-1)
                picked = m[pos]
                assert (picked[:-1] >= picked[1:]).all()
        if prev is not None:
            assert ((prev == 1) & (flat == 0)).sum().item() == 0            # nested across thresholds
        prev = flat


def test_masked_adamw_bit_exact():
    from multimodal_unlearning_b200 import salun
    dev = torch.device("cuda:0")
    n = 1_000_003  # not a multiple of 8: exercises the tail
    g = torch.Generator().manual_seed(3)
    p = torch.randn(n, generator=g).to(torch.bfloat16)
    mask = (torch.rand(n, generator=g) < 0.5).to(torch.int8)
    m = np.zeros(n, np.float32)
    v = np.zeros(n, np.float32)
    pd = p.clone().to(dev)
    md = torch.zeros(n, dtype=torch.float32, device=dev)
    vd = torch.zeros(n, dtype=torch.float32, device=dev)
    for step in (1, 2, 3):
        grad = (0.01 * torch.randn(n, generator=g)).to(torch.bfloat16)
        p, m, v = SO.adamw_step(p, grad, m, v, mask.numpy(), step, lr=1e-3, weight_decay=0.01)
        salun.masked_adamw_flat(pd, grad.to(dev), md, vd, mask.to(dev), step, lr=1e-3, weight_decay=0.01)
        assert torch.equal(pd.cpu().view(torch.int16), p.view(torch.int16)), step
        assert np.array_equal(md.cpu().numpy().view(np.int32), m.view(np.int32)), step
        assert np.array_equal(vd.cpu().numpy().view(np.int32), v.view(np.int32)), step


def test_masked_adamw_optimizer_matches_torch_adamw_when_mask_is_ones():
    from multimodal_unlearning_b200 import salun
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    w = torch.randn(256, 128, device=dev).to(torch.bfloat16)
    p1 = torch.nn.Parameter(w.clone())
    p2 = torch.nn.Parameter(w.float().clone())
    o1 = salun.MaskedAdamW([p1], lr=1e-3, weight_decay=0.01)
    o2 = torch.optim.AdamW([p2], lr=1e-3, weight_decay=0.01, foreach=False)
    o1.set_mask(p1, torch.ones_like(w, dtype=torch.int8))
    for _ in range(3):
        g = (0.01 * torch.randn(256, 128, device=dev)).to(torch.bfloat16)
        p1.grad = g.clone()
        p2.grad = g.float()
        o1.step()
        o2.step()
        p2.data = p2.data.to(torch.bfloat16).float()  # our parameters are stored in bf16 after every step
    assert (p1.float() - p2).abs().max().item() <= 2 ** -7 * p2.abs().max().item()
    assert torch.allclose(o1.state[p1]["exp_avg"], o2.state[p2]["exp_avg"], rtol=1e-5, atol=1e-9)


# ---- bit-packed masks -------------------------------------------------------------
def test_packed_masks_equal_packed_int8_masks():
    """clear_saliency_mask_bits == np.packbits(clear_saliency_mask), on the tie-heavy / special-value tensors too."""
    from multimodal_unlearning_b200 import salun
    g = torch.Generator().manual_seed(0)
    named = {"a": ((torch.randn(1000003, generator=g) * 4).round() / 4), "b": torch.randn(8, generator=g),
             "c": torch.randn(3, 5, generator=g), "d": torch.randn(2048, 1024, generator=g), "e": torch.randn(1, generator=g)}
    named["d"][:, ::3] = 0.0
    named = {n: v.to(torch.bfloat16) for n, v in named.items()}
    layout = salun.FlatLayout([(n, v.shape) for n, v in named.items()])
    state = salun.SaliencyState(layout, "cuda:0")
    state.accumulate_flat(_flat(layout, named, torch.bfloat16).cuda())
    ths = (0.0, 0.1, 0.5, 0.999, 1.0)
    sel = state.select(ths)
    for j, th in enumerate(ths):
        m8 = state.mask_flat(sel, j).cpu().numpy()
        bits = state.mask_flat(sel, j, packed=True)
        assert bits.dtype == torch.uint8 and bits.numel() == layout.total // 8
        assert np.array_equal(bits.cpu().numpy(), np.packbits(m8.astype(np.uint8), bitorder="little")), th
        # and the per-tensor slices are byte aligned: tensor i starts at byte offsets[i] / 8
        want = SO.masks(named, (th,))[th]
        for i, n in enumerate(layout.names):
            got = salun.unpack_mask(bits[layout.offsets[i] // 8: layout.offsets[i] // 8 + (layout.numels[i] + 7) // 8],
                                    named[n].shape)
            assert np.array_equal(got.cpu().numpy(), want[n]), (th, n)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 64, 100_003])
def test_pack_unpack_round_trip(n):
    from multimodal_unlearning_b200 import salun
    g = torch.Generator().manual_seed(n)
    m = (torch.rand(n, generator=g) < 0.4).to(torch.int8)
    bits = salun.pack_mask(m.cuda())
    assert np.array_equal(bits.cpu().numpy(), np.packbits(m.numpy().astype(np.uint8), bitorder="little"))
    assert torch.equal(salun.unpack_mask(bits, (n,)).cpu(), m)


def test_masked_adamw_packed_mask_bit_exact():
    """The update driven by the bit-packed mask equals the int8-mask update and the oracle, bit for bit."""
    from multimodal_unlearning_b200 import salun
    dev = torch.device("cuda:0")
    n = 1_000_003
    g = torch.Generator().manual_seed(4)
    p = torch.randn(n, generator=g).to(torch.bfloat16)
    mask = (torch.rand(n, generator=g) < 0.5).to(torch.int8)
    grad = (0.01 * torch.randn(n, generator=g)).to(torch.bfloat16)
    want_p, want_m, want_v = SO.adamw_step(p, grad, np.zeros(n, np.float32), np.zeros(n, np.float32), mask.numpy(), 1, lr=1e-3)
    pd, md, vd = p.to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    salun.masked_adamw_flat(pd, grad.to(dev), md, vd, salun.pack_mask(mask.to(dev)), 1, lr=1e-3)
    assert torch.equal(pd.cpu().view(torch.int16), want_p.view(torch.int16))
    assert np.array_equal(md.cpu().numpy().view(np.int32), want_m.view(np.int32))
    assert np.array_equal(vd.cpu().numpy().view(np.int32), want_v.view(np.int32))


def test_mask_files_round_trip_into_the_optimizer(tmp_path):
    """save_gradient_ratio writes the reference's `with_{th}.pt` (dict name -> int8 tensor, generate_mask.py:72-89);
    load_masks / set_masks_from_file read it back bit-packed; the packed update equals the int8 one."""
    import types
    from multimodal_unlearning_b200 import norms, salun
    dev = torch.device("cuda:0")
    torch.manual_seed(0)

    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.emb = torch.nn.Embedding(50, 16)
            self.out = torch.nn.Linear(16, 50, bias=False)

        def forward(self, input_ids, attention_mask=None, labels=None):
            z = self.out(self.emb(input_ids)).float()
            return types.SimpleNamespace(loss=torch.nn.functional.cross_entropy(z[:, :-1].reshape(-1, 50), labels[:, 1:].reshape(-1)))

    model = Tiny().to(dev).to(torch.bfloat16)
    ids = torch.randint(0, 50, (2, 9))
    salun.save_gradient_ratio([(ids, ids, torch.ones_like(ids))], model, None, types.SimpleNamespace(save_dir=str(tmp_path)))
    raw = torch.load(tmp_path / "with_0.5.pt")
    assert set(raw) == {n for n, _ in model.named_parameters()}
    for n, p in model.named_parameters():
        assert raw[n].dtype == torch.int8 and raw[n].shape == p.shape and int(raw[n].sum()) == int(0.5 * p.numel())
    packed = salun.load_masks(str(tmp_path / "with_0.5.pt"), device=dev)
    for n, p in model.named_parameters():
        assert packed[n].dtype == torch.uint8
        assert torch.equal(salun.unpack_mask(packed[n], p.shape).cpu(), raw[n])
    # two optimizers: masks from the file (packed) vs the same masks as int8; clipping active in both
    import copy
    m2 = copy.deepcopy(model)
    o1 = salun.MaskedAdamW(model.parameters(), lr=1e-2, max_grad_norm=0.05)
    o1.set_masks_from_file(str(tmp_path / "with_0.5.pt"), model)
    o2 = salun.MaskedAdamW(m2.parameters(), lr=1e-2, max_grad_norm=0.05)
    for (n, p) in m2.named_parameters():
        o2.set_mask(p, raw[n].to(dev), packed=False)
    for _ in range(2):
        for pa, pb in zip(model.parameters(), m2.parameters()):
            gr = torch.randn(pa.shape, device=dev).to(torch.bfloat16)
            pa.grad, pb.grad = gr.clone(), gr.clone()
        o1.step()
        o2.step()
        assert o1.last_grad_norm.item() == o2.last_grad_norm.item()
    for (n, pa), pb in zip(model.named_parameters(), m2.parameters()):
        assert torch.equal(pa, pb), n
        frozen = raw[n].to(dev) == 0
        assert frozen.any() and not torch.equal(pa, torch.zeros_like(pa))


def test_histogram_over_every_bf16_pattern():
    """The histogram is taken over bf16 magnitude patterns and mapped to fp16 keys at flush time: every one of the 65536
    bf16 bit patterns except the NaNs (two tensors, so that the lane-private window sits at different places), against
    np.bincount of the oracle's keys."""
    import ctypes as C
    from multimodal_unlearning_b200 import _lib, salun
    dev = torch.device("cuda:0")
    pats = torch.arange(0x10000, dtype=torch.int32)
    finite_or_inf = (pats & 0x7FFF) <= 0x7F80
    allp = pats[finite_or_inf].to(torch.int16).view(torch.bfloat16)
    g = torch.Generator().manual_seed(9)
    bulk = (1e-3 * torch.randn(200_000, generator=g)).to(torch.bfloat16)
    named = {"all": torch.cat([allp, allp, bulk]), "tiny": torch.cat([allp, (1e-30 * torch.randn(50_000, generator=g)).to(torch.bfloat16)])}
    layout = salun.FlatLayout([(n, v.shape) for n, v in named.items()])
    state = salun.SaliencyState(layout, dev)
    state.accumulate_flat(_flat(layout, named, torch.bfloat16).to(dev))
    state.select((0.5,))
    hist = state.hist.cpu().numpy()
    for i, n in enumerate(layout.names):
        want = np.bincount(SO.saliency_keys(named[n]), minlength=32768)
        pad = layout.offsets[i + 1] - layout.offsets[i] - layout.numels[i]     # zero padding of the flat layout counts as key 0
        want[0] += pad
        assert np.array_equal(hist[i].astype(np.int64), want), n


def test_full_llama7b_parameter_list_properties():
    """BASELINE configs[3] size: all 291 tensors / 6.74 B parameters of Llama-2-7B.  Far beyond the CPU sort, so
    order-statistic properties: exactly int(th * numel) ones per tensor, masks nested across thresholds, the bit-packed
    mask equal to the packed int8 mask, and the masked update touching exactly the selected parameters."""
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1] / "tools"))
    import bench_salun
    from multimodal_unlearning_b200 import salun
    dev = torch.device("cuda:0")
    if torch.cuda.get_device_properties(0).total_memory < 100e9:
        pytest.skip("needs the B200's 180 GB")
    torch.cuda.empty_cache()   # ~95 GB at the peak of this test: start from an empty allocator cache
    lay = salun.FlatLayout(bench_salun.llama7b_shapes(32))
    assert len(lay.names) == 291 and lay.total == 6738415616
    st = salun.SaliencyState(lay, dev)
    g = torch.Generator(device=dev).manual_seed(0)
    for o in range(0, lay.total, 1 << 28):
        n = min(1 << 28, lay.total - o)
        st.acc[o:o + n] = (1e-3 * torch.randn(n, generator=g, device=dev)).to(torch.bfloat16)
    ths = (0.1, 0.5)
    sel = st.select(ths)
    m10 = st.mask_flat(sel, 0)
    m50 = st.mask_flat(sel, 1)
    assert int(((m10 == 1) & (m50 == 0)).sum().item()) == 0                      # nested
    numels = torch.tensor(lay.numels, dtype=torch.int64)
    offs = lay.offsets
    for m, th in ((m10, 0.1), (m50, 0.5)):
        per8 = m.view(-1, 8).sum(1, dtype=torch.int32)                                    # ones per 8-element group
        csum = torch.cumsum(per8, 0, dtype=torch.int64)                                   # ones up to every group boundary
        del per8
        ends = torch.tensor([(offs[i] + (lay.numels[i] + 7) // 8 * 8) // 8 - 1 for i in range(len(lay.names))], device=dev)
        starts = torch.tensor([offs[i] // 8 - 1 for i in range(len(lay.names))], device=dev)
        upto = csum[ends]
        before = torch.where(starts >= 0, csum[starts.clamp(min=0)], torch.zeros_like(upto))
        counts = (upto - before).cpu()
        want = torch.tensor([int(th * n) for n in lay.numels], dtype=torch.int64)
        assert torch.equal(counts, want), th
        del csum
    bits = st.mask_flat(sel, 1, packed=True)
    assert torch.equal(salun.pack_mask(m50), bits)             # converter kernel vs direct bit-packed output
    step = 1 << 27                                               # and against torch, slice by slice
    weights = (2 ** torch.arange(8, device=dev, dtype=torch.int16)).view(1, 8)
    for o in range(0, lay.total, step):
        sl = m50[o:o + step].view(-1, 8).to(torch.int16)
        assert torch.equal((sl * weights).sum(1).to(torch.uint8), bits[o // 8:(o + step) // 8])
    del m10
    # masked update: parameters outside the mask only see the weight decay, parameters inside move
    p = torch.ones(lay.total, dtype=torch.bfloat16, device=dev)
    mom = torch.zeros(lay.total, dtype=torch.float32, device=dev)
    var = torch.zeros(lay.total, dtype=torch.float32, device=dev)
    salun.masked_adamw_flat(p, st.acc, mom, var, bits, 1, lr=1e-2, weight_decay=0.0)
    moved = p != 1
    sel_nonzero = (m50 == 1) & (st.acc != 0)
    assert torch.equal(moved, sel_nonzero)
    assert int((mom != 0).sum().item()) == int(sel_nonzero.sum().item())
    del p, mom, var, moved, sel_nonzero, m50, bits, st
    torch.cuda.empty_cache()


def _check_global(named, ths):
    from multimodal_unlearning_b200 import salun
    layout = salun.FlatLayout([(n, v.shape) for n, v in named.items()])
    state = salun.SaliencyState(layout, "cuda:0")
    state.accumulate_flat(_flat(layout, named, torch.bfloat16).cuda())
    want = SO.masks_global(named, ths)
    got = dict(state.masks(ths, scope="global"))
    for th in ths:
        for n in named:
            g = got[th][n].cpu().numpy()
            assert np.array_equal(g, want[th][n]), ("global", th, n, int((g != want[th][n]).sum()))
    return state, layout


def test_global_scope_bit_exact():
    """scope="global": ONE ranking over all tensors (k = int(th * total), ties by global flat index) — bit-exact against the
    oracle's stable sort of the concatenated keys.  Tensors of very different scales, heavy ties that straddle tensors,
    exact zeros next to the zero padding between tensors, odd sizes, fp16 overflow, th = 0 and th = 1."""
    g = torch.Generator().manual_seed(4)
    named = {"a": (torch.randn(1000003, generator=g) * 4).round() / 4,        # heavy ties
             "b": torch.randn(8, generator=g), "c": torch.randn(3, 5, generator=g),
             "d": (torch.randn(2048, 1024, generator=g) * 4).round() / 4,     # the same tie values again, in a later tensor
             "e": torch.randn(1, generator=g), "f": torch.randn(70001, generator=g) * 1e-9,   # flushes to fp16 zero
             "z": torch.zeros(1003), "h": torch.randn(333, 77, generator=g) * 1e-3}
    named["d"][:, ::3] = 0.0
    named["d"][5, 7] = 1e6                                                    # overflows fp16 -> inf
    named = {n: v.to(torch.bfloat16) for n, v in named.items()}
    ths = (0.0, 0.05, 0.1, 0.25, 0.5, 0.75, 0.9, 0.999, 1.0)
    state, layout = _check_global(named, ths)
    # the reference's ten thresholds as well, and the bit-packed format
    sel = state.select(SO.REFERENCE_THRESHOLDS, scope="global")
    want = SO.masks_global(named, SO.REFERENCE_THRESHOLDS)
    total = sum(layout.numels)
    from multimodal_unlearning_b200 import salun
    for j, th in enumerate(SO.REFERENCE_THRESHOLDS):
        bits = state.mask_flat(sel, j, packed=True)
        ones = 0
        for i, n in enumerate(layout.names):
            nb = (layout.numels[i] + 7) // 8
            m = salun.unpack_mask(bits[layout.offsets[i] // 8: layout.offsets[i] // 8 + nb].clone(), layout.shapes[i]).cpu().numpy()
            assert np.array_equal(m, want[th][n]), (th, n)
            ones += int(m.sum())
        assert ones == int(th * total)


def test_global_scope_single_tensor_equals_per_tensor():
    g = torch.Generator().manual_seed(5)
    named = {"only": ((torch.randn(50000, generator=g) * 8).round() / 8).to(torch.bfloat16)}
    state, _ = _check_global(named, (0.1, 0.5, 0.9))
    per = dict(state.masks((0.1, 0.5, 0.9)))
    glo = dict(state.masks((0.1, 0.5, 0.9), scope="global"))
    for th in per:
        assert torch.equal(per[th]["only"], glo[th]["only"])


# --------------------------------------------------------------------------------------------------
# 8-bit block-wise optimiser states — parity unpinned upstream (bitsandbytes absent): bit-exact against
# the oracle's restatement of the scheme, and bounded against the fp32-state update
# --------------------------------------------------------------------------------------------------
def test_adamw8_codebooks():
    from multimodal_unlearning_b200 import salun
    for signed in (True, False):
        code = salun.dynamic_codebook(signed).numpy()
        want = SO.dynamic_map(signed)
        assert code.shape == (256,) and np.all(np.diff(code) > 0) and code[-1] == 1.0
        assert (code.min() == 0.0) != signed and np.count_nonzero(code == 0.0) == 1
        # same published construction; linspace arithmetic (torch fp32 vs numpy) may differ in the last bit
        assert np.max(np.abs(code - want) / np.maximum(np.abs(want), 1e-30)) <= 2e-7
        lut = salun.quantisation_tables(torch.from_numpy(code), signed)
        assert lut.numel() == (65536 if signed else 32768) and lut.dtype == torch.uint8


@pytest.mark.parametrize("n", [4096, 70001, 256 * 300 + 8])
@pytest.mark.parametrize("mask_kind", ["none", "int8", "bits"])
def test_adamw8_bit_exact_vs_oracle(n, mask_kind):
    """Several steps of clear_masked_adamw_8bit against oracle.adamw8_step: parameters (bf16 bits), both code-index arrays and
    both absmax arrays identical after every step; masked-out parameters keep zero moments; an all-zero block stays exactly zero."""
    from multimodal_unlearning_b200 import norms, salun
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(n)
    p = torch.randn(n, generator=g).to(torch.bfloat16)
    mask = (torch.rand(n, generator=g) < 0.5).to(torch.int8).numpy() if mask_kind != "none" else None
    c1, c2 = salun.dynamic_codebook(True).numpy(), salun.dynamic_codebook(False).numpy()
    s1, s2, a1, a2 = salun.adamw8_init(n, dev)
    o1, o2, oa1, oa2 = s1.cpu().numpy(), s2.cpu().numpy(), a1.cpu().numpy(), a2.cpu().numpy()
    pd, po = p.to(dev), p.clone()
    md = None
    if mask is not None:
        md = torch.from_numpy(mask).to(dev)
        if mask_kind == "bits":
            md = salun.pack_mask(md)
    hyper = dict(lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.01)
    for step in range(1, 5):
        gr = (torch.randn(n, generator=g) * (0.0 if step == 3 else 10.0 ** torch.randint(-4, 1, (1,), generator=g).item())).to(torch.bfloat16)
        gr[:512] = 0                                               # two blocks that never see a gradient
        scale = None
        if step == 4:                                              # fused clipping coefficient on the last step
            sumsq = norms.grad_sqnorm([gr.to(dev)], [md])
            scale = norms.clip_coef(sumsq, 0.5)
        salun.masked_adamw8_flat(pd, gr.to(dev), s1, s2, a1, a2, md, step, lr=1e-3, grad_scale=scale)
        po, o1, o2, oa1, oa2 = SO.adamw8_step(po, gr, o1, o2, oa1, oa2, mask, step, code1=c1, code2=c2,
                                              grad_scale=float(scale.item()) if scale is not None else None, **hyper)
        torch.cuda.synchronize()
        assert torch.equal(pd.cpu().view(torch.int16), po.view(torch.int16)), step
        assert np.array_equal(s1.cpu().numpy(), o1) and np.array_equal(s2.cpu().numpy(), o2), step
        assert np.array_equal(a1.cpu().numpy(), oa1) and np.array_equal(a2.cpu().numpy(), oa2), step
    assert a1[:2].abs().sum().item() == 0 and a2[:2].abs().sum().item() == 0
    if mask is not None:
        zero1, zero2 = int(np.nonzero(c1 == 0)[0][0]), int(np.nonzero(c2 == 0)[0][0])
        off = mask == 0
        assert np.all(s1.cpu().numpy()[off] == zero1) and np.all(s2.cpu().numpy()[off] == zero2)


def test_adamw8_tracks_fp32_states_and_saves_memory():
    """MaskedAdamW(state_bits=8) against the fp32-state optimiser over 30 steps on the same gradients: the parameters stay
    within a small multiple of bf16 resolution of each other; optimiser state is 2 bytes + 8/256 per parameter instead of 8."""
    from multimodal_unlearning_b200 import salun
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    shapes = [(512, 384), (4096,), (100,)]                          # the 100-element tensor keeps fp32 moments
    base = [torch.randn(*s, generator=g).to(torch.bfloat16) for s in shapes]
    p32 = [torch.nn.Parameter(b.clone().to(dev)) for b in base]
    p8 = [torch.nn.Parameter(b.clone().to(dev)) for b in base]
    o32 = salun.MaskedAdamW(p32, lr=1e-3, weight_decay=0.0)
    o8 = salun.MaskedAdamW(p8, lr=1e-3, weight_decay=0.0, state_bits=8)
    for _ in range(30):
        for a, b in zip(p32, p8):
            gr = (0.1 * torch.randn(a.shape, generator=g)).to(torch.bfloat16).to(dev)
            a.grad, b.grad = gr, gr.clone()
        o32.step()
        o8.step()
    for a, b, b0 in zip(p32, p8, base):
        moved = (a.detach().float().cpu() - b0.float()).abs().mean().item()
        diff = (a.detach().float() - b.detach().float()).abs().mean().item()
        assert moved > 0 and diff <= 0.08 * moved, (diff, moved)    # the 8-bit trajectory follows the fp32 one
    st = o8.state[p8[0]]
    assert st["state1"].dtype == torch.uint8 and st["state1"].numel() == p8[0].numel() and "exp_avg" not in st
    assert "exp_avg" in o8.state[p8[2]]
    bytes8 = sum(t.numel() * t.element_size() for k, t in st.items() if torch.is_tensor(t))
    assert bytes8 <= (2 + 8 / 256) * p8[0].numel() + 16
