"""Stamp pyproject.toml with a nightly version ``X.Y.Z.devYYYYMMDD`` (reference: scripts/patch_nightly_version.py)."""

import datetime
import pathlib
import re

path = pathlib.Path(__file__).resolve().parent.parent / "pyproject.toml"
text = path.read_text()
m = re.search(r'^version = "([0-9]+\.[0-9]+\.[0-9]+)[^"]*"', text, flags=re.M)
assert m, "no version line in pyproject.toml"
stamp = datetime.datetime.now(datetime.timezone.utc).strftime("%Y%m%d")
path.write_text(text.replace(m.group(0), f'version = "{m.group(1)}.dev{stamp}"'))
print(f"{m.group(1)}.dev{stamp}")
