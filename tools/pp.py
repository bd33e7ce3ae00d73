#!/usr/bin/env python
"""Prints selected keys of the JSON line that tools/profile_step.py writes (stdin)."""
import json
import sys

d = json.loads(sys.stdin.read())
keys = sys.argv[2:] or ["device_ms", "ms_extend", "ms_shade", "ms_shadow", "ms_film", "mean"]
print(sys.argv[1] if len(sys.argv) > 1 else "", {k: round(d[k], 3) for k in keys if k in d})
