"""`python -m kungfu_b200.cmd -np 4 python train.py` == `kungfu-run -np 4 python train.py` (for installs without the console script)."""
from . import run

run()
