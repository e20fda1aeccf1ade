"""Paired input pipeline + epoch metrics.

CPU: the oracle (oracle/augment_ref.py) against the fixtures the reference's own transform lists produced
(tests/golden/augment_ref.pt, oracle/gen_golden.py:gen_augment) and against torchvision itself; the host logic of
mfvit.data (random draws, fixed-point coefficients, sharding, AUC).  GPU: the kernels, through the C ABI, bit-exact
against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import augment_ref as A


def _golden(golden_dir):
    return torch.load(os.path.join(golden_dir, "augment_ref.pt"))


def test_oracle_matches_reference_transform_fixtures(golden_dir):
    cases = _golden(golden_dir)
    assert len(cases) == 6
    for c in cases:
        img = c["img"].numpy()
        mean, std = A.STATS[c["img_type"]]
        torch.manual_seed(c["seed"])
        got = A.transform_train(img, c["crop"], c["rotate"], mean, std)
        assert torch.equal(got, c["train"]), (c["img_type"], c["rotate"])        # bit-exact, float32
        assert torch.equal(A.transform_eval(img, c["crop"], mean, std), c["eval"])


def test_oracle_matches_torchvision_and_consumes_the_same_random_numbers():
    T = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    rng = np.random.default_rng(5)
    for trial in range(24):
        h, w, crop = [(224, 224, 224), (256, 256, 224), (240, 256, 224), (64, 80, 48)][trial % 4]
        deg = [0, 1, True, 10, False, 30][trial % 6]
        img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
        mean, std = A.STATS["Train_Mix" if trial % 2 else "data"]
        tf = T.Compose([T.RandomHorizontalFlip(), T.RandomRotation(deg), T.RandomCrop((crop, crop)), T.ToTensor(),
                        T.Normalize(mean=mean, std=std)])
        torch.manual_seed(trial)
        ref = tf(Image.fromarray(img))
        after_ref = torch.rand(1)
        torch.manual_seed(trial)
        got = A.transform_train(img, crop, deg, mean, std)
        after_got = torch.rand(1)
        assert torch.equal(ref, got), trial
        assert torch.equal(after_ref, after_got)  # same number of draws -> the next sample sees the same stream


def test_host_draws_and_coefficients_follow_the_oracle():
    from mfvit import data
    for h, w, crop, deg in [(224, 224, 224, True), (256, 256, 224, 1), (72, 80, 64, 10), (224, 224, 224, False)]:
        gen = torch.Generator()
        gen.manual_seed(1234)
        ours = data.draw_train_params(9, h, w, crop, deg, gen)
        torch.manual_seed(1234)  # same Mersenne stream as the seeded generator
        theirs = [A.draw_train(h, w, crop, deg) for _ in range(9)]
        assert ours == theirs
        packed = data.pack_params(ours, w, h)
        assert packed.dtype == torch.int32 and tuple(packed.shape) == (9, data.N_PARAMS)
        for row, (flip, angle, top, left) in zip(packed.tolist(), theirs):
            fixed = A.rotation_fixed(angle, w, h)
            assert row[0] == int(flip) and row[8] == top and row[9] == left
            assert row[1] == (0 if fixed is None else 1)
            assert tuple(row[2:8]) == (fixed if fixed is not None else (0,) * 6)
    gen = torch.Generator()
    gen.manual_seed(77)
    ours = data.draw_train_params_batch(33, 72, 80, 64, 5, gen)
    torch.manual_seed(77)
    assert ours == A.draw_train_batch(33, 72, 80, 64, 5)
    assert {f for f, _, _, _ in ours} == {True, False} and all(-5 <= a <= 5 and 0 <= t <= 8 and 0 <= l <= 16 for _, a, t, l in ours)
    assert data.eval_params(2, 256, 240, 224) == [(False, 0.0, *A.center_crop_offsets(256, 240, 224))] * 2
    assert data.STATS == A.STATS
    with pytest.raises(data.MfvError):
        data.rotation_coefficients(180.0, 64, 64)


def test_shards_are_aligned_disjoint_and_equal_sized():
    from mfvit import data
    n = 103
    full = data.shard_indices(n, epoch=3, seed=7, shuffle=True)
    assert sorted(full.tolist()) == list(range(n))
    assert not torch.equal(full, data.shard_indices(n, epoch=4, seed=7, shuffle=True))
    parts = [data.shard_indices(n, 3, 7, True, rank=r, world_size=4) for r in range(4)]
    assert len({len(p) for p in parts}) == 1 and len(parts[0]) == 26            # padded by wrap-around, like DistributedSampler
    assert set(torch.cat(parts).tolist()) == set(range(n))
    dropped = [data.shard_indices(n, 3, 7, True, rank=r, world_size=4, drop_last=True) for r in range(4)]
    assert sum(len(p) for p in dropped) == 100 and len(set(torch.cat(dropped).tolist())) == 100
    assert torch.equal(data.shard_indices(5, 0, 0, False), torch.arange(5))


def test_store_validates_and_loader_refuses_cpu():
    from mfvit import data
    u8 = torch.zeros(4, 8, 8, 3, dtype=torch.uint8)
    store = data.PairedU8Store(u8, u8.clone(), [0, 1, 2, 1], pin=False)
    assert len(store) == 4 and store.labels.dtype == torch.int64
    with pytest.raises(data.MfvError):
        data.PairedU8Store(u8, u8[:3], [0, 1, 2, 1], pin=False)
    with pytest.raises(data.MfvError):
        data.PairedU8Store(u8.float(), u8, [0, 1, 2, 1], pin=False)
    with pytest.raises(data.MfvError):
        data.PairedDeviceLoader(store, 2, crop=8, device="cpu")
    for shape in [(40, 24, 24, 3), (10, 5, 7, 3)]:  # row bytes divisible by 8 (word gather) and not (byte gather)
        st = torch.randint(0, 256, shape, dtype=torch.uint8)
        idx = torch.randperm(shape[0])[:4]
        dst = torch.zeros(6, *shape[1:], dtype=torch.uint8)
        data._gather_rows(st, idx, dst[:4])
        assert torch.equal(dst[:4], st[idx]) and int(dst[4:].sum()) == 0


def test_auc_matches_pairwise_definition_and_sklearn():
    from mfvit import data
    rng = np.random.default_rng(3)
    vals = np.round(rng.normal(size=(200, 3)), 1).astype(np.float32)  # rounding -> ties
    gts = rng.integers(0, 3, size=200)
    acc, auc = A.epoch_metrics(vals, gts)
    assert abs(data.roc_auc_ovr_mean(vals, gts, 3) - auc) < 1e-12
    metrics = pytest.importorskip("sklearn.metrics")
    from sklearn.preprocessing import label_binarize
    onehot = label_binarize(gts, classes=[0, 1, 2])  # MAIN_CA:897-903
    ref = []
    for c in range(3):
        fpr, tpr, _ = metrics.roc_curve(onehot[:, c], vals[:, c])
        ref.append(metrics.auc(fpr, tpr))
    assert abs(float(np.mean(ref)) - auc) < 1e-9


# ---------------------------------------------------------------------------------------------------------------- GPU
def _kernel_vs_oracle(img_batch, samples, crop, img_type):
    from mfvit import data, ops
    B, H, W, _ = img_batch.shape
    mean, std = A.STATS[img_type]
    dev = "cuda"
    params = data.pack_params(samples, W, H).to(dev)
    out = ops.augment_u8(torch.from_numpy(img_batch).to(dev), params, torch.tensor(mean, device=dev),
                         torch.tensor(std, device=dev), crop)
    torch.cuda.synchronize()
    for i, (flip, angle, top, left) in enumerate(samples):
        want = A.to_tensor_normalize(A.apply_u8(img_batch[i], flip, A.rotation_fixed(angle, W, H), top, left, crop), mean, std)
        assert torch.equal(out[i].cpu(), want), (i, flip, angle, top, left)
    return out


@pytest.mark.gpu
def test_augment_kernel_is_bit_exact(golden_dir):
    rng = np.random.default_rng(11)
    for H, W, crop, img_type in [(224, 224, 224, "data"), (256, 256, 224, "Train_Mix"), (240, 272, 224, "CheXpert_Enh"),
                                 (64, 80, 48, "CheXpert-v1.0-small"), (400, 400, 384, "data")]:
        B = 5
        imgs = rng.integers(0, 256, size=(B, H, W, 3), dtype=np.uint8)
        samples = [(False, 0.0, 0, 0), (True, 0.0, H - crop, W - crop), (False, 1.0, (H - crop) // 2, 0),
                   (True, -0.73, 0, (W - crop) // 2), (True, 29.5, (H - crop) // 3, (W - crop) // 3)]
        _kernel_vs_oracle(imgs, samples, crop, img_type)
    # and straight against what the reference's transform lists produced
    from mfvit import data, ops
    for c in _golden(golden_dir):
        img = c["img"].numpy()
        h, w, _ = img.shape
        torch.manual_seed(c["seed"])
        s = A.draw_train(h, w, c["crop"], c["rotate"])
        out = _kernel_vs_oracle(img[None], [s], c["crop"], c["img_type"])
        assert torch.equal(out[0].cpu(), c["train"])
        out = _kernel_vs_oracle(img[None], data.eval_params(1, h, w, c["crop"]), c["crop"], c["img_type"])
        assert torch.equal(out[0].cpu(), c["eval"])
    with pytest.raises(data.MfvError):
        ops.augment_u8(torch.zeros(1, 8, 8, 3, dtype=torch.uint8, device="cuda"), torch.zeros(1, 12, dtype=torch.int32, device="cuda"),
                       torch.zeros(3, device="cuda"), torch.ones(3, device="cuda"), 6)  # crop % 4
    # a crop window that leaves the image (bad device-side parameters) reads nothing out of bounds: fill value 0
    img = rng.integers(0, 256, size=(1, 32, 32, 3), dtype=np.uint8)
    mean, std = A.STATS["Train_Mix"]
    out = ops.augment_u8(torch.from_numpy(img).cuda(), data.pack_params([(True, 0.0, 8, 8)], 32, 32).cuda(),
                         torch.tensor(mean, device="cuda"), torch.tensor(std, device="cuda"), 32)
    padded = np.zeros((40, 40, 3), dtype=np.uint8)
    padded[:32, :32] = img[0][:, ::-1]
    assert torch.equal(out[0].cpu(), A.to_tensor_normalize(padded[8:40, 8:40], mean, std))


@pytest.mark.gpu
def test_paired_loader_yields_aligned_transformed_pairs():
    from mfvit import data
    rng = np.random.default_rng(21)
    N, H, W, crop, B = 23, 72, 80, 64, 8
    cxr = rng.integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)
    enh = rng.integers(0, 256, size=(N, H, W, 3), dtype=np.uint8)
    labels = rng.integers(0, 3, size=N)
    store = data.PairedU8Store(cxr, enh, labels)
    seed, epoch = 5, 2
    loader = data.PairedDeviceLoader(store, B, crop=crop, degrees=True, training=True, seed=seed)
    loader.set_epoch(epoch)
    assert len(loader) == 3
    idx = data.shard_indices(N, epoch, seed, True)
    torch.manual_seed(seed * 1000003 + epoch * 1009)  # replay the loader's generator through the oracle
    seen = []
    for bi, (xc, xe, y) in enumerate(loader):
        ids = idx[bi * B:(bi + 1) * B].tolist()
        assert xc.shape == (len(ids), 3, crop, crop) and xc.is_cuda and y.dtype == torch.int64
        assert y.cpu().tolist() == [int(labels[i]) for i in ids]
        xc, xe = xc.cpu(), xe.cpu()
        for src, got, t in ((cxr, xc, "data"), (enh, xe, "Train_Mix")):
            mean, std = A.STATS[t]
            drawn = A.draw_train_batch(len(ids), H, W, crop, True)
            for j, i in enumerate(ids):
                flip, angle, top, left = drawn[j]
                want = A.to_tensor_normalize(A.apply_u8(src[i], flip, A.rotation_fixed(angle, W, H), top, left, crop), mean, std)
                assert torch.equal(got[j], want), (bi, j, t)
        seen += ids
    assert sorted(seen) == list(range(N))  # one pass, every pair once, both types indexed by the same permutation
    # eval mode: centre crop, no randomness, natural order
    ev = data.PairedDeviceLoader(store, B, crop=crop, training=False, shuffle=False)
    for xc, xe, y in ev:
        pass  # 23 = 8 + 8 + 7: the last batch is ragged
    assert xc.shape[0] == 7 and y.cpu().tolist() == [int(v) for v in labels[16:]]
    mean, std = A.STATS["Train_Mix"]
    assert torch.equal(xe[3].cpu(), A.transform_eval(enh[19], crop, mean, std))


@pytest.mark.gpu
def test_epoch_metrics_accumulate_on_device():
    from mfvit import data
    rng = np.random.default_rng(8)
    m = data.EpochMetrics(capacity=100, num_classes=3)
    all_vals, all_gt, loss_sum = [], [], 0.0
    for rows in (32, 32, 7):
        a, b, c = (torch.from_numpy(rng.normal(size=(rows, 3)).astype(np.float32)).cuda() for _ in range(3))
        a[0, :] = 0.25
        b[0, :] = 0.5
        c[0, :] = -0.75  # an all-equal row: torch.max picks index 0
        t = torch.from_numpy(rng.integers(0, 3, size=rows)).cuda()
        loss = torch.tensor([float(rng.uniform(0.5, 1.5))], device="cuda")
        m.accumulate(a, b, c, t, loss)
        all_vals.append(((a + b) + c).cpu().numpy())
        all_gt.append(t.cpu().numpy())
        loss_sum += float(loss.item()) * rows
    vals, gts = np.concatenate(all_vals), np.concatenate(all_gt)
    ep_loss, ep_auc, ep_acc = m.result()
    acc, auc = A.epoch_metrics(vals, gts)
    assert np.array_equal(m.vals[:71].cpu().numpy(), vals)                                  # bit-exact scores
    assert np.array_equal(m.preds[:71].cpu().numpy(), torch.from_numpy(vals).max(1)[1].numpy())
    assert ep_acc == acc and abs(ep_auc - auc) < 1e-12 and abs(ep_loss - loss_sum / 71) < 1e-9
    m.reset()
    small = data.EpochMetrics(capacity=4, num_classes=3)
    small.accumulate(a, None, None, t, loss)
    with pytest.raises(data.MfvError):
        small.result()


@pytest.mark.gpu
def test_loader_feeds_the_captured_step_and_metrics_stay_on_device():
    """uint8 store -> PairedDeviceLoader -> MFViTCATrainer.step (captured graph) with EpochMetrics:
    the same losses as the step fed with the oracle-transformed float32 tensors, and the epoch's loss / accuracy / AUC
    equal to the per-step host computation of MAIN_CA:884-909."""
    import e2e_common as E
    from mfvit import data
    from mfvit.trainer import MFViTCATrainer
    rng = np.random.default_rng(31)
    N, B, crop = 16, 8, 224
    cxr = rng.integers(0, 256, size=(N, crop, crop, 3), dtype=np.uint8)
    enh = rng.integers(0, 256, size=(N, crop, crop, 3), dtype=np.uint8)
    labels = rng.integers(0, 3, size=N)
    store = data.PairedU8Store(cxr, enh, labels)
    runs = []
    for use_loader in (True, False):
        _, (o_f, o_c, o_e) = E.build_mfvit_pair(seed=9)
        metrics = data.EpochMetrics(capacity=N, num_classes=3)
        tr = MFViTCATrainer(o_f, o_c, o_e, lr=1e-3, momentum=0.9, metrics=metrics, train_backbones=True)
        tr.capture_graph(*E.synthetic_pair(B, crop, device="cuda"))
        assert int(metrics.counters[0]) == 0  # warm-up steps of the capture are not part of the epoch
        losses, vals = [], []
        if use_loader:
            loader = data.PairedDeviceLoader(store, B, crop=crop, degrees=True, training=True, seed=3)
            for xc, xe, y in loader:
                losses.append(float(tr.step(xc, xe, y)))
                f, a, b = tr.logits()
                vals.append(((f + a) + b).cpu().numpy())
            assert tr.graph_replays == 2
        else:
            idx = data.shard_indices(N, 0, 3, True)
            torch.manual_seed(3 * 1000003)
            for bi in range(2):
                ids = idx[bi * B:(bi + 1) * B].tolist()
                xs = []
                for src, t in ((cxr, "data"), (enh, "Train_Mix")):
                    mean, std = A.STATS[t]
                    drawn = A.draw_train_batch(B, crop, crop, crop, True)
                    xs.append(torch.stack([A.to_tensor_normalize(
                        A.apply_u8(src[i], f, A.rotation_fixed(ang, crop, crop), tp, lf, crop), mean, std)
                        for i, (f, ang, tp, lf) in zip(ids, drawn)]).cuda())
                y = torch.from_numpy(labels[ids]).cuda()
                losses.append(float(tr.step(xs[0], xs[1], y)))
                f, a, b = tr.logits()
                vals.append(((f + a) + b).cpu().numpy())
        ep_loss, ep_auc, ep_acc = metrics.result()
        idx = data.shard_indices(N, 0, 3, True).numpy()
        acc, auc = A.epoch_metrics(np.concatenate(vals), labels[idx])
        assert ep_acc == acc and (abs(ep_auc - auc) < 1e-12 or (np.isnan(auc) and np.isnan(ep_auc)))
        assert abs(ep_loss - sum(l * B for l in losses) / N) < 1e-6
        runs.append(losses)
    for a, b in zip(*runs):
        assert abs(a - b) <= 1e-3 * max(1.0, abs(a)), runs  # identical inputs; fp32 reduce-add order may differ


@pytest.mark.gpu
def test_run_phase_val_matches_the_oracle_and_train_updates():
    """mfvit.loops.run_phase: the `val` phase (forward only, MAIN_CA:824-909 with phase == 'val') reproduces the oracle's
    loss / accuracy / AUC on the centre-cropped images and leaves the parameters alone; the `train` phase steps them."""
    import torch.nn.functional as F
    import e2e_common as E
    from mfvit import data, loops
    from mfvit.trainer import MFViTCATrainer
    rng = np.random.default_rng(41)
    N, B, H, crop = 12, 8, 240, 224
    cxr = rng.integers(0, 256, size=(N, H, H, 3), dtype=np.uint8)
    enh = rng.integers(0, 256, size=(N, H, H, 3), dtype=np.uint8)
    labels = rng.integers(0, 3, size=N)
    store = data.PairedU8Store(cxr, enh, labels)
    (r_f, r_c, r_e), (o_f, o_c, o_e) = E.build_mfvit_pair(seed=17)
    metrics = data.EpochMetrics(capacity=N, num_classes=3)
    tr = MFViTCATrainer(o_f, o_c, o_e, lr=1e-3, momentum=0.9, metrics=metrics, train_backbones=True)
    val = data.PairedDeviceLoader(store, B, crop=crop, training=False, shuffle=False)
    v_loss, v_auc, v_acc = loops.run_phase("val", tr, val, metrics, N)
    master0 = tr.engine.master.clone()
    # oracle: eval transform + as-written fp32 graph
    outs, loss_sum = [], 0.0
    with torch.no_grad():
        for lo in range(0, N, B):
            ids = range(lo, min(lo + B, N))
            xc = torch.stack([A.transform_eval(cxr[i], crop, *A.STATS["data"]) for i in ids]).cuda()
            xe = torch.stack([A.transform_eval(enh[i], crop, *A.STATS["Train_Mix"]) for i in ids]).cuda()
            fused, x_c, x_e = r_f(r_c, r_e, xc, xe)
            out = fused + x_c + x_e
            t = torch.from_numpy(labels[lo:lo + B]).cuda()
            loss_sum += float(F.cross_entropy(out, t)) * len(ids)
            outs.append(out.cpu().numpy())
    ref_vals = np.concatenate(outs)
    assert np.abs(metrics.vals[:N].cpu().numpy() - ref_vals).max() <= 3 * 2e-3   # three logit sets, 2e-3 each (north_star)
    assert abs(v_loss - loss_sum / N) <= 5e-3
    acc, auc = A.epoch_metrics(metrics.vals[:N].cpu().numpy(), labels)
    assert v_acc == acc and abs(v_auc - auc) < 1e-12
    again = loops.run_phase("val", tr, val, metrics, N)
    assert again == (v_loss, v_auc, v_acc) and torch.equal(tr.engine.master, master0)
    train = data.PairedDeviceLoader(store, B, crop=crop, degrees=True, training=True, seed=1, drop_last=True)
    t_loss, _, t_acc = loops.run_phase("train", tr, train, metrics, B)
    assert np.isfinite(t_loss) and not torch.equal(tr.engine.master, master0)
    assert int(metrics.counters[0]) == B
    with pytest.raises(data.MfvError):
        loops.run_phase("train", MFViTCATrainer(o_f, o_c, o_e, train_backbones=True), train, metrics)


def test_store_from_csv_follows_the_reference_dataset(tmp_path):
    """PairedU8Store.from_csv: list format and decode path of loader.py:Dataset_covid (fields[1]/folder/fields[2], label =
    fields[-2], cv2.imread -> PIL) and Resize((img_size, img_size)) of image_transform.py:55, for both image types."""
    cv2 = pytest.importorskip("cv2")
    T = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    from mfvit import data
    rng = np.random.default_rng(2)
    root = str(tmp_path)
    names, labels = ["a.png", "b.png", "c.png"], [2, 0, 1]
    for folder in ("data", "Train_Mix"):
        os.makedirs(os.path.join(root, folder))
        for n in names:
            cv2.imwrite(os.path.join(root, folder, n), rng.integers(0, 256, size=(40, 52, 3), dtype=np.uint8))
    csv = os.path.join(root, "train.txt")
    with open(csv, "w") as f:
        for i, (n, l) in enumerate(zip(names, labels)):
            f.write("%d %s %s %d x\n" % (i, root, n, l))
    store = data.PairedU8Store.from_csv("data", "Train_Mix", csv, img_size=32)
    assert len(store) == 3 and store.labels.tolist() == labels and tuple(store.cxr.shape) == (3, 32, 32, 3)
    for folder, got in (("data", store.cxr), ("Train_Mix", store.enh)):
        for i, n in enumerate(names):
            want = np.asarray(T.Resize((32, 32))(Image.fromarray(cv2.imread(os.path.join(root, folder, n)))))
            assert np.array_equal(got[i].numpy(), want)
    with pytest.raises(data.MfvError):
        data.PairedU8Store.from_csv("data", "Train_Mix", csv, img_size=32, maintain_ratio=True)
