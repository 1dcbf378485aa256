#!/usr/bin/env python
"""Input pipeline study: ImageNet-style shards through the elastic data adaptor, with the augmentation on the device.

Reference: experimental/gradient_noise/prepro.py (TFRecords -> `ExampleDatasetAdaptor` -> `adaptive_train`: per epoch, update
the offset, the topology and rewind). Here: pre-decoded uint8 shards (`helpers.imagenet.write_shard`), the `ElasticDataAdaptor`
(global sample offset + (np, rank) sharding that survives a resize), batches pinned on the host and augmented on the GPU
(`helpers.imagenet.train_augment`) when one is present.

    python experimental/gradient_noise/prepro.py --make-synthetic 2 --data-dir /tmp/kf-imagenet
    kungfu-run -np 2 python experimental/gradient_noise/prepro.py --data-dir /tmp/kf-imagenet --adaptive
"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import kungfu_b200.torch as kf  # noqa: E402
from kungfu_b200.elastic.datasets import ElasticDataAdaptor  # noqa: E402
from kungfu_b200.helpers import imagenet, show_rate  # noqa: E402


def parse():
    ap = argparse.ArgumentParser(description="imagenet dataset example", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    ap.add_argument("--data-dir", default=os.path.join(os.path.expanduser("~"), "Desktop"), help="dir to dataset")
    ap.add_argument("--batch-size", type=int, default=32, help="input batch size")
    ap.add_argument("--pattern", default="train", help="subset prefix of the shard names (<subset>-XXXXX-of-YYYYY.npz)")
    ap.add_argument("--adaptive", action="store_true", help="enable adaptive training (offset + topology updates per epoch)")
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--make-synthetic", type=int, default=0, metavar="SHARDS", help="write SHARDS synthetic shards and exit")
    ap.add_argument("--size", type=int, default=64, help="output crop (synthetic source images are 1.25x this)")
    return ap.parse_args()


def make_synthetic(a):
    os.makedirs(a.data_dir, exist_ok=True)
    rng = np.random.RandomState(0)
    src = int(a.size * 1.25)
    for i in range(a.make_synthetic):
        imagenet.write_shard(os.path.join(a.data_dir, f"{a.pattern}-{i:05d}-of-{a.make_synthetic:05d}.npz"),
                             rng.randint(0, 256, (128, src, src, 3), dtype=np.uint8), rng.randint(0, 1000, 128))
    print("wrote", a.make_synthetic, "shards to", a.data_dir)


def simple_train(a, dev):
    n_img, t0 = 0, time.time()
    shard = (kf.current_rank(), kf.current_cluster_size())
    for epoch in range(a.epochs):
        for x, y in imagenet.create_dataset_from_files(imagenet.list_shards(a.data_dir, a.pattern), a.batch_size, shard=shard):
            img = imagenet.train_augment(x.to(dev, non_blocking=True), (a.size, a.size), dtype=torch.float32)
            n_img += img.shape[0]
    return n_img, time.time() - t0, img


def adaptive_train(a, dev):
    """The data is one global sequence; each worker reads its stride, and the offset is re-agreed every epoch (after a resize a
    newcomer would adopt the survivors' offset through the same max all-reduce)."""
    xs, ys = [], []
    for name in imagenet.list_shards(a.data_dir, a.pattern):
        with np.load(name) as z:
            xs.append(z["images"])
            ys.append(z["labels"])
    data, labels = torch.from_numpy(np.concatenate(xs)), torch.from_numpy(np.concatenate(ys))
    ds = ElasticDataAdaptor(data, labels, a.batch_size)
    n_img, t0 = 0, time.time()
    steps_per_epoch = max(1, len(data) // (a.batch_size * kf.current_cluster_size()))
    for epoch in range(a.epochs):
        offset = int(kf.ops.all_reduce_fn(torch.tensor([ds.offset]), op="max").item())  # update_offset
        ds.set_offset(offset)
        if kf.current_rank() == 0:
            print(f"BEGIN epoch {epoch}: offset {offset}, np {kf.current_cluster_size()}", flush=True)
        for _ in range(steps_per_epoch):
            x, y = ds.next_batch()
            img = imagenet.train_augment(x.to(dev, non_blocking=True), (a.size, a.size), dtype=torch.float32)
            n_img += img.shape[0]
        steps_per_epoch = max(1, len(data) // (a.batch_size * kf.current_cluster_size()))  # update_topology
    return n_img, time.time() - t0, img


def main():
    a = parse()
    if a.make_synthetic:
        return make_synthetic(a)
    dev = "cpu"
    if torch.cuda.is_available() and kf.get_cuda_index() >= 0:
        torch.cuda.set_device(kf.get_cuda_index())
        dev = "cuda"
    n_img, dt, img = (adaptive_train if a.adaptive else simple_train)(a, dev)
    if dev == "cuda":
        torch.cuda.synchronize()
    total = int(kf.ops.all_reduce_fn(torch.tensor([n_img])).item())
    if kf.current_rank() == 0:
        print(f"RESULT: {total} images in {dt:.2f}s = {total / dt:.1f} img/s ({show_rate(total * img[0].numel() * img.element_size(), dt)} of "
              f"augmented tensors), last batch {tuple(img.shape)} mean {img.float().mean().item():+.3f}", flush=True)


if __name__ == "__main__":
    main()
