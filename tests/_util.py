import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def level_text(n: int) -> str:
    with open(os.path.join(GOLDEN, "levels", f"lvl{n}")) as f:
        return f.read()
