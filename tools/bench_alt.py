#!/usr/bin/env python
"""Run bench.py against another build of the library (A/B timing on the GPU box): ALT_LIB=/path/to/lib.so
python tools/bench_alt.py <bench.py arguments>.  Only this wrapper redirects the loader; the package itself
always loads the in-tree libsbce.so."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sbce  # noqa: E402

sbce._lib.LIB_PATH = os.environ["ALT_LIB"]
import bench  # noqa: E402

sys.exit(bench.main())
