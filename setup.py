"""pip install -e . — builds the native pieces in-tree (no Go, no TensorFlow; the reference needs both:
setup_tensorflow.py:70-123, setup_pytorch.py:21-53) and registers the `kungfu-run` console script."""
from setuptools import find_packages, setup
from setuptools.command.build_py import build_py


class BuildNative(build_py):
    def run(self):
        import os
        import sys

        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from kungfu_b200 import _build

        _build.build_all(verbose=True)
        super().run()


setup(
    name="kungfu_b200",
    version="0.1.0",
    description="Blackwell-native adaptive distributed training for PyTorch (KungFu capabilities, B200-first design)",
    packages=find_packages(include=["kungfu_b200", "kungfu_b200.*", "kungfu", "kungfu.*"]),
    package_data={"kungfu_b200": ["*.so", "bin/*", "lib/*", "csrc/**/*"]},
    cmdclass={"build_py": BuildNative},
    entry_points={"console_scripts": ["kungfu-run=kungfu_b200.cmd:run", "kungfu-rrun=kungfu_b200.cmd.rrun:main",
                                      "kungfu-test-util=kungfu_b200.cmd.test_util:main",
                                      "kungfu-config-client=kungfu_b200.cmd.config_client:main"]},
    python_requires=">=3.10",
)
