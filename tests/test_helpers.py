"""Dataset helpers, experimental tools and shell launch helpers (reference: srcs/python/kungfu/tensorflow/v1/helpers/*,
experimental/{atari_pingpong,gradient_noise}, scripts/launcher.sh)."""
import os
import pickle
import subprocess
import sys

import numpy as np
import pytest
import torch

from test_runtime_multiproc import ROOT, kungfu_run, ports

from kungfu_b200 import helpers
from kungfu_b200.helpers import cifar, idx, imagenet, mnist


@pytest.mark.parametrize("dtype", [np.uint8, np.int8, np.int16, np.int32, np.float32, np.float64])
@pytest.mark.parametrize("gz", [False, True])
def test_idx_roundtrip(tmp_path, dtype, gz):
    a = (np.random.RandomState(0).randn(3, 4, 5) * 50).astype(dtype)
    path = str(tmp_path / ("a.idx.gz" if gz else "a.idx"))
    idx.write_idx_file(path, a)
    b = idx.read_idx_file(path)
    assert b.dtype == a.dtype and b.shape == a.shape and (a == b).all()
    if not gz:  # header layout: 0 0 <code> <rank> then big-endian dims
        raw = open(path, "rb").read()
        assert raw[:2] == b"\0\0" and raw[3] == 3 and raw[4:16] == bytes([0, 0, 0, 3, 0, 0, 0, 4, 0, 0, 0, 5])


def test_idx_rejects_bad_input(tmp_path):
    p = tmp_path / "bad.idx"
    p.write_bytes(b"\x01\x00\x08\x01\x00\x00\x00\x01\x00")
    with pytest.raises(ValueError):
        idx.read_idx_file(str(p))
    p.write_bytes(b"\x00\x00\x08\x01\x00\x00\x00\x09\x00")  # promises 9 bytes, has 1
    with pytest.raises(ValueError):
        idx.read_idx_file(str(p))
    with pytest.raises(ValueError):
        idx.write_idx_file(str(tmp_path / "c.idx"), np.zeros(3, np.complex64))


def test_npz_to_idx_tar(tmp_path):
    src = str(tmp_path / "w.npz")
    np.savez(src, a=np.arange(6, dtype=np.float32).reshape(2, 3), b=np.arange(4, dtype=np.int32))
    tar = idx.npz2idxtar(src)
    assert tar.endswith("w.idx.tar")
    back = idx.idxtar2dict(tar)
    assert set(back) == {"a", "b"} and (back["a"] == np.arange(6).reshape(2, 3)).all() and back["b"].dtype == np.int32


def test_mnist_loader(tmp_path):
    sets = mnist.synthetic(100, 20, seed=3)
    mnist.write_datasets(str(tmp_path), sets)
    raw = mnist.load_datasets(str(tmp_path))
    assert raw.train.images.shape == (100, 28, 28) and raw.train.images.dtype == np.uint8 and raw.test.labels.shape == (20,)
    assert (raw.train.images == sets.train.images).all() and (raw.test.labels == sets.test.labels).all()
    full = mnist.load_datasets(str(tmp_path), normalize=True, one_hot=True, padded=True)
    assert full.train.images.shape == (100, 32, 32, 1) and full.train.images.dtype == np.float32 and full.train.images.max() <= 1.0
    assert full.train.labels.shape == (100, 10) and (full.train.labels.sum(1) == 1).all()
    assert (full.train.labels.argmax(1) == raw.train.labels).all()
    with pytest.raises(ValueError):
        mnist.load_mnist_data(str(tmp_path), "valid", False, False)


def test_cifar_loaders(tmp_path):
    rng = np.random.RandomState(0)
    d10 = tmp_path / "cifar-10-batches-py"
    d10.mkdir()
    for name in [f"data_batch_{i}" for i in range(1, 6)] + ["test_batch"]:
        with open(d10 / name, "wb") as f:
            pickle.dump({b"data": rng.randint(0, 256, (20, 3072), dtype=np.uint8), b"labels": [int(v) for v in rng.randint(0, 10, 20)]}, f)
    sets = cifar.Cifar10Loader(str(tmp_path), normalize=True, one_hot=True).load_datasets()
    assert sets.train.images.shape == (100, 32, 32, 3) and sets.train.images.dtype == np.float32 and sets.train.labels.shape == (100, 10)
    assert sets.test.images.shape == (20, 32, 32, 3)
    d100 = tmp_path / "cifar-100-python"
    d100.mkdir()
    for name, n in (("train", 30), ("test", 10)):
        with open(d100 / name, "wb") as f:
            pickle.dump({b"data": rng.randint(0, 256, (n, 3072), dtype=np.uint8), b"fine_labels": [int(v) for v in rng.randint(0, 100, n)]}, f)
    s100 = cifar.Cifar100Loader(str(tmp_path)).load_datasets()
    assert s100.train.images.shape == (30, 32, 32, 3) and s100.train.images.dtype == np.uint8 and s100.test.labels.max() < 100
    # the raw layout is planar RGB: pixel (0,0) of image 0 is bytes 0, 1024, 2048
    first = pickle.load(open(d100 / "train", "rb"), encoding="bytes")[b"data"][0]
    assert tuple(s100.train.images[0, 0, 0]) == (first[0], first[1024], first[2048])


def test_imagenet_pipeline(tmp_path):
    rng = np.random.RandomState(0)
    for i in range(2):
        imagenet.write_shard(str(tmp_path / f"train-{i:05d}-of-00002.npz"), rng.randint(0, 256, (10, 40, 48, 3), dtype=np.uint8),
                             np.arange(10 * i, 10 * i + 10))
    batches = list(imagenet.create_dataset(str(tmp_path), batch_size=8, n=2))
    assert len(batches) == 2 and batches[0][0].shape == (8, 40, 48, 3) and batches[1][1].tolist() == list(range(8, 16))
    assert len(list(imagenet.create_dataset(str(tmp_path), batch_size=8, n=1))) == 1  # n limits the shard count
    r0 = [y for _, y in imagenet.create_dataset(str(tmp_path), batch_size=5, n=2, shard=(0, 2))]
    r1 = [y for _, y in imagenet.create_dataset(str(tmp_path), batch_size=5, n=2, shard=(1, 2))]
    assert sorted(torch.cat(r0 + r1).tolist()) == list(range(20))  # the two workers cover every sample once
    g = torch.Generator().manual_seed(0)
    x = imagenet.train_augment(batches[0][0], out_hw=(32, 32), dtype=torch.float32, generator=g)
    assert x.shape == (8, 3, 32, 32) and x.is_contiguous(memory_format=torch.channels_last) and torch.isfinite(x).all()
    g2 = torch.Generator().manual_seed(0)
    assert torch.equal(x, imagenet.train_augment(batches[0][0], out_hw=(32, 32), dtype=torch.float32, generator=g2))
    # with the full-image crop, no flip randomness left out and no colour jitter, a constant image stays constant
    const = torch.full((2, 40, 48, 3), 128, dtype=torch.uint8)
    y = imagenet.train_augment(const, out_hw=(16, 16), color=False, dtype=torch.float32)
    want = torch.tensor([(128 / 255 - m) / s for m, s in zip(imagenet.MEAN, imagenet.STD)]).view(1, 3, 1, 1)
    assert torch.allclose(y, want.expand_as(y), atol=1e-5)
    e = imagenet.eval_preprocess(const, out_hw=(16, 16), dtype=torch.float32)
    assert torch.allclose(e, want.expand_as(e), atol=1e-5)
    img, label = imagenet.record_to_labeled_image({"image": const[0].numpy(), "label": 7}, training=False)
    assert img.shape == (3, 224, 224) and label == 7


def test_random_input_and_units():
    x, y = helpers.random_input((4, 3), 10)
    assert x.shape == (4, 3) and 0 <= int(y) < 9
    assert helpers.show_size(1536) == "1.50Ki" and helpers.show_size(12) == "12" and helpers.show_rate(1 << 30, 0.5) == "2.00GiB/s"


def test_plot_noise_parses_launcher_logs(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "experimental", "gradient_noise"))
    import plot_noise

    log = tmp_path / "run.log"
    log.write_text("\n".join([
        '\x1b[1;32m[127.0.0.1.10000::stdout]\x1b[m {"phase_np": 4, "img_per_s": 100.0, "noise_scale": 60.0}',
        '[127.0.0.1.10001::stdout] {"noise_scale": 1.0}',
        'RESULT: 5.0 {"noise_scale": 14.0, "loss": 2.5}',
        "not json {",
    ]))
    s = plot_noise.parse_log(str(log))
    assert [v for _, v in s["noise_scale"]] == [60.0, 1.0, 14.0] and [v for _, v in s["loss"]] == [2.5]
    only = plot_noise.parse_log(str(log), worker="10000")
    assert [v for _, v in only["noise_scale"]] == [60.0, 14.0]  # unprefixed lines are kept, other workers dropped
    assert plot_noise.ema([1.0, 3.0], 0.5) == [1.0, 2.0] and len(plot_noise.sparkline(list(range(500)))) <= 64
    assert plot_noise.main([str(log), "--out", str(tmp_path / "n"), "--ema", "0.5"]) == 0
    assert (tmp_path / "n.csv").read_text().splitlines()[0] == "index,noise_scale,loss,img_per_s"


def test_pingpong_data_parallel_policy_gradient(tmp_path):
    script = os.path.join(ROOT, "experimental", "pingpong", "pingpong.py")
    ckpt = str(tmp_path / "pp.npz")
    env = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_CONFIG_LOG_LEVEL="WARN")
    r = subprocess.run([sys.executable, script, "--init", "--checkpoint", ckpt], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                       timeout=120)
    assert r.returncode == 0 and os.path.exists(ckpt), r.stdout[-2000:]
    rport, prange = ports()
    r = subprocess.run([kungfu_run(), "-q", "-np", "2", "-H", "127.0.0.1:2", "-port", str(rport), "-port-range", prange, sys.executable, script,
                        "--checkpoint", ckpt, "--batch-size", "1", "--episodes", "3"], env=env, cwd=ROOT, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=240)
    assert r.returncode == 0 and "replicas identical = True" in r.stdout, r.stdout[-3000:]


def test_launcher_sh_helpers(tmp_path):
    rport, prange = ports()
    script = tmp_path / "t.sh"
    script.write_text(f"""
KF_LAUNCHER_SH={ROOT}/scripts/launcher.sh
. {ROOT}/scripts/launcher.sh
export KF_LOGDIR={tmp_path}/logs
srun echo single-ok
prun 2 -H 127.0.0.1:2 -port {rport} -port-range {prange} {sys.executable} {ROOT}/examples/torch_simple_example.py
""")
    env = dict(os.environ, PYTHONPATH=ROOT, KUNGFU_RUN=kungfu_run(), KUNGFU_CONFIG_LOG_LEVEL="WARN")
    r = subprocess.run(["sh", str(script)], env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=180)
    assert r.returncode == 0 and "single-ok" in r.stdout and "rank 1/2" in r.stdout and "BEGIN" in r.stdout and "END" in r.stdout, r.stdout[-3000:]
    assert os.path.isdir(tmp_path / "logs")


def test_vision_model_zoo_matches_reference_shapes():
    """The three models of the reference's scalability charts (benchmarks/system/benchmark_kungfu.py:58-66)."""
    from kungfu_b200 import models
    from kungfu_b200.benchmarks import model_sizes

    vgg16 = models.build("vgg16")
    sizes = [p.numel() for p in vgg16.parameters()]
    assert sum(sizes) == 138_357_544 and len(sizes) == 32 and sizes == model_sizes.VGG16  # reference model_sizes.py:1-5
    inc = models.build("inception_v3")
    assert sum(p.numel() for p in inc.parameters()) == 23_834_568 and models.input_size("inception_v3") == 299
    assert sum(p.numel() for p in models.build("inception_v3", aux_logits=True).parameters()) == 27_161_264
    assert sum(model_sizes.sizes("inception_v3")) == 23_834_568 and sum(model_sizes.sizes("resnet50")) == 25_557_032
    small = models.vgg.vgg11(num_classes=10).eval()  # eval: no dropout noise
    x = torch.randn(2, 3, 64, 64)
    a = small(x)
    b = small(x.contiguous(memory_format=torch.channels_last))
    assert a.shape == (2, 10) and torch.allclose(a, b, atol=1e-5), "VGG output must not depend on the memory format"
    inc.train()
    out = inc(torch.randn(2, 3, 299, 299))
    assert out.shape == (2, 1000)
    out.sum().backward()
    assert all(p.grad is not None for p in inc.parameters())
    with pytest.raises(ValueError):
        models.build("alexnet")


def test_reference_import_names_resolve():
    """A script written against the reference keeps its imports: `import kungfu.torch as kf`, `kungfu.python`, `kungfu.cmd`."""
    import kungfu
    import kungfu.python as kfp
    import kungfu.torch as kf
    from kungfu._utils import measure, show_duration
    from kungfu.cmd import launch_multiprocess  # noqa: F401
    from kungfu.torch.ops import all_reduce_fn, broadcast_parameters, inplace_all_reduce_op  # noqa: F401
    from kungfu.torch.optimizers import (PairAveragingOptimizer, SynchronousAveragingOptimizer,  # noqa: F401
                                         SynchronousSGDOptimizer)

    import kungfu_b200.torch

    assert kf is kungfu_b200.torch and kungfu.__version__
    assert kfp.current_rank() == 0 and kfp.current_cluster_size() == 1 and kfp.detached() is False
    assert torch.equal(all_reduce_fn(torch.ones(3)), torch.ones(3))
    dt, val = measure(lambda: 7)
    assert val == 7 and dt >= 0 and isinstance(show_duration(1.5), str)
    m = torch.nn.Linear(3, 2)
    opt = SynchronousSGDOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), named_parameters=m.named_parameters())
    assert isinstance(opt, torch.optim.SGD)  # still an instance of the user's optimizer class
    m(torch.ones(1, 3)).sum().backward()
    opt.step()
    assert isinstance(kungfu.info.collect(), dict) and kfp._resize_from_url is not None


def test_global_variables_registry_and_dataset_adaptors():
    """Reference: tensorflow/variables.py (named global variables, batch-size / noise-scale accessors, setters) and
    v1/datasets/adaptor.py (BaseDatasetAdaptor / ExampleDatasetAdaptor)."""
    from kungfu_b200 import variables as kv
    from kungfu_b200.elastic import BaseDatasetAdaptor, ExampleDatasetAdaptor, PolicyHook, ScheduledBatchSizePolicy

    for n in list(kv.global_variables()):
        kv.remove_global_variable(n)
    assert kv.get_global_variable("x") is None and kv.get_batch_size() is None
    with pytest.raises(RuntimeError):
        kv.eval_global_variable("x")
    v = kv.create_global_variable("x", shape=(3,), dtype=torch.float32)
    assert v.shape == (3,) and not v.requires_grad and kv.get_or_create_global_variable("x") is v
    with pytest.raises(ValueError):
        kv.create_global_variable("x")
    set_x = kv.create_setter(v)
    set_x([1.0, 2.0, 3.0])
    assert kv.eval_global_variable("x").tolist() == [1.0, 2.0, 3.0]
    op, place = kv.create_assign_op_for(v)
    place.fill_(7.0)
    op()
    assert v.tolist() == [7.0, 7.0, 7.0] and kv.create_placeholder_for(v).shape == v.shape
    bs = kv.get_or_create_batch_size(init=32)
    assert bs.dtype == torch.int32 and kv.eval_batch_size() == 32 and kv.get_batch_size() is bs
    kv.create_setter(bs)(48)
    assert kv.eval_batch_size() == 48
    kv.remove_global_variable(kv.GraphKeys.BATCH_SIZE)
    # the policy hook publishes its counters; a policy that changes the batch size is visible through the registry
    hook = PolicyHook([ScheduledBatchSizePolicy({0: 10, 1: 20})], epoch_size=100, epoch_num=2, init_batch_size=5)
    hook.begin()
    seen = []
    while hook.before_step():
        seen.append(kv.eval_batch_size())
        hook.after_step()
    hook.end()
    assert seen == [10] * 10 + [20] * 5 and int(kv.eval_global_variable(kv.GraphKeys.TRAINED_SAMPLES)) == 200
    assert int(kv.eval_global_variable(kv.GraphKeys.TOTAL_SAMPLES)) == 200
    # the noise-scale monitor owns kungfu_gradient_noise_scale
    m = torch.nn.Linear(2, 1)
    import kungfu_b200.torch as kf

    kf.optimizers.MonitorGradientNoiseScaleOptimizer(torch.optim.SGD(m.parameters(), lr=0.1), m.named_parameters(), device_batch_size=4)
    assert kv.eval_gradient_noise_scale() == 0.0
    kv.sync_global_variables()  # single process: a no-op that must not fail
    # dataset adaptors
    data = torch.arange(40)
    a = BaseDatasetAdaptor(batch_size=4, shard_count=2, shard_id=1)
    batches = [b.tolist() for b in a(data)]
    assert batches[0] == [4, 5, 6, 7] and batches[1] == [12, 13, 14, 15] and len(batches) == 5
    assert a.create_update_offset()() == 8 and [b.tolist() for b in a(data)][0] == [12, 13, 14, 15]
    assert a.create_rewind()() == 0 and "offset=0, np=2, rank=1, batch size=4" == a.debug()
    with pytest.raises(RuntimeError):
        a.create_update_topology()
    e = ExampleDatasetAdaptor(batch_size=4, shard_count=2, shard_id=1)
    e.create_update_topology()()
    assert (e._shard_count, e._shard_id, e._batch_size) == (3, 2, 6)


def test_image_trainer_host_path_matches_a_manual_step():
    """`ImageTrainer.step()` without a GPU: uint8 NHWC batch -> normalise -> forward/backward -> optimizer, same numbers as the
    hand-written loop (the GPU path replays the same body from a CUDA graph)."""
    import kungfu_b200.torch as kf
    from kungfu_b200.trainer import IMAGENET_MEAN, IMAGENET_STD, ImageTrainer

    torch.manual_seed(0)
    mk = lambda: torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(),  # noqa: E731
                                     torch.nn.Linear(4, 5))
    a, b = mk(), mk()
    b.load_state_dict(a.state_dict())
    opt = kf.optimizers.SynchronousSGDOptimizer(torch.optim.SGD(a.parameters(), lr=0.05, momentum=0.9), named_parameters=a.named_parameters())
    tr = ImageTrainer(a, opt, 6, 16, torch.float32, use_graph=True, device=torch.device("cpu"))  # use_graph is ignored on the host
    ref = torch.optim.SGD(b.parameters(), lr=0.05, momentum=0.9)
    g = torch.Generator().manual_seed(1)
    for _ in range(3):
        x = torch.randint(0, 256, (6, 16, 16, 3), dtype=torch.uint8, generator=g)
        y = torch.randint(0, 5, (6,), generator=g)
        loss = tr.step(x, y)
        xn = ((x.float() - torch.tensor(IMAGENET_MEAN)) / torch.tensor(IMAGENET_STD)).permute(0, 3, 1, 2)
        ref.zero_grad()
        want = torch.nn.functional.cross_entropy(b(xn), y)
        want.backward()
        ref.step()
        assert abs(loss - float(want.detach())) < 1e-5
    for p, q in zip(a.parameters(), b.parameters()):
        assert torch.allclose(p, q, atol=1e-6)
    assert tr.h2d_bytes_per_step == 0 and tr.d2h_bytes_per_step == 4
