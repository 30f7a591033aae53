"""Regenerates the data fixtures that come from the reference tree (run in the build container only;
/root/reference does not exist on the GPU box).

  * tests/golden/levels/lvl1..6 and lle_b200/resources/levels/lvl1..6 : the six built-in maps
    (reference: resources/levels/lvl1..6, embedded by src/core/levels.rs:1-8).  They are *data* the
    drop-in must ship to honour ``World.level(n)``.
  * tests/golden/layouts.json : the map strings of the reference's layout corpus
    (python/tests/world_layouts.py, 34 named layouts), used as a differential-fuzz corpus.

Usage: python tests/golden/make_fixtures.py [/root/reference]
"""
import json
import os
import sys
import types

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))

for dst in (os.path.join(HERE, "levels"), os.path.join(ROOT, "lle_b200", "resources", "levels")):
    os.makedirs(dst, exist_ok=True)
    for n in range(1, 7):
        with open(os.path.join(REF, "resources", "levels", f"lvl{n}")) as f:
            text = f.read()
        with open(os.path.join(dst, f"lvl{n}"), "w") as f:
            f.write(text)

# world_layouts.py only needs `lle.World` to exist at import time.
stub = types.ModuleType("lle")
stub.World = object
sys.modules["lle"] = stub
sys.path.insert(0, os.path.join(REF, "python", "tests"))
import world_layouts  # noqa: E402

layouts = {}
for layout in world_layouts.ALL_LAYOUTS:
    if isinstance(layout.source, int):
        with open(os.path.join(REF, "resources", "levels", f"lvl{layout.source}")) as f:
            layouts[layout.name] = f.read()
    else:
        layouts[layout.name] = layout.source
with open(os.path.join(HERE, "layouts.json"), "w") as f:
    json.dump(layouts, f, indent=1, sort_keys=True)
print(f"{len(layouts)} layouts, 6 levels written")
