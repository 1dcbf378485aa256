"""Three tiny programs with known answers: identity, all-reduce of ones, and S-SGD on f(x) = x^2 (every worker computes the same
gradient, so the averaged update is the single-worker one: x_k = (1 - 2 lr)^k). Extra here: the same closed form through the
model-averaging optimizers, on the GPU kernels when the workers own GPUs."""
import sys

import torch

import kungfu_b200.torch as kf


def device():
    if torch.cuda.is_available() and kf.get_cuda_index() >= 0:
        torch.cuda.set_device(kf.get_cuda_index())
        return torch.device("cuda", kf.get_cuda_index())
    return torch.device("cpu")


def show_info_example():
    print("rank=%d, np=%d" % (kf.current_rank(), kf.current_cluster_size()), flush=True)


def all_reduce_example(dev):
    x = torch.ones((), dtype=torch.int32, device=dev)
    for step in range(5):
        v = int(kf.ops.all_reduce_fn(x))
        print("step %d, result: %d" % (step, v), flush=True)
        if v != kf.current_cluster_size():
            raise RuntimeError("unexpected all-reduce result %d" % v)


def sgd_example(dev, make, label, exact=True):
    lr = 0.1
    x = torch.nn.Parameter(torch.ones(8, device=dev))
    opt = make(torch.optim.SGD([x], lr=lr), [("x", x)])
    for step in range(5):
        opt.zero_grad()
        (x * x).sum().backward()
        opt.step()
        v = float(x.detach()[0])
        u = (1 - 2 * lr) ** (step + 1)
        print("%s step %d, result: %f" % (label, step, v), flush=True)
        if not exact:  # asynchronous gossip: a peer may be a step ahead or behind, only the envelope is known
            if not (0.0 < v <= (1 - 2 * lr) ** max(0, step - 3) + 1e-6):
                raise RuntimeError("%s left the envelope: %f" % (label, v))
        elif abs(u - v) > 1e-5:
            msg = "%s unexpected result: %f, want: %f" % (label, v, u)
            print(msg, flush=True)
            raise RuntimeError(msg)


def main(args):
    dev = device()
    show_info_example()
    all_reduce_example(dev)
    sgd_example(dev, lambda o, n: kf.optimizers.SynchronousSGDOptimizer(o, named_parameters=n), "S-SGD")
    # identical replicas stay identical under synchronous model averaging, so the closed form holds there too
    sgd_example(dev, lambda o, n: kf.optimizers.SynchronousAveragingOptimizer(o, named_parameters=n), "SMA")
    sgd_example(dev, lambda o, n: kf.optimizers.PairAveragingOptimizer(o, named_parameters=n), "PairAveraging", exact=False)
    kf.run_barrier()
    if kf.current_rank() == 0:
        print("examples ok", flush=True)


if __name__ == "__main__":
    main(sys.argv)
