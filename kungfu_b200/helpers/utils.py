"""Size / rate pretty-printers (reference: helpers/utils.py:6-26)."""


def show_size(s: float) -> str:
    """Binary units with two decimals: 1536 -> '1.50Ki'."""
    units = ["", "Ki", "Mi", "Gi", "Ti"]
    v = float(s)
    for u in units:
        if abs(v) < 1024 or u == units[-1]:
            return f"{v:.2f}{u}" if u else f"{int(v)}"
        v /= 1024
    return f"{v:.2f}Ti"


def show_rate(size: float, duration: float) -> str:
    return show_size(size / max(duration, 1e-12)) + "B/s"
