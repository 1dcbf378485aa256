"""Model zoo used by the benchmarks and tests: ResNet (headline), VGG and Inception v3 (the other two models of the
reference's scalability charts), BERT (PairAveraging), SLP (plumbing)."""
from . import bert, inception, resnet, slp, vgg  # noqa: F401


def build(name: str, **kw):
    """`build("resnet50")`, `build("vgg16")`, `build("inception_v3")`, ... — the benchmark scripts' model switch."""
    for mod in (resnet, vgg, inception):
        f = getattr(mod, name, None)
        if callable(f):
            return f(**kw)
    raise ValueError(f"unknown vision model {name!r}")


def input_size(name: str) -> int:
    """Side of the square ImageNet crop the model is benchmarked on (keras-applications defaults)."""
    return 299 if name.startswith("inception") else 224
