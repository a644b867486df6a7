"""profile_step.py with the per-op path forced (use_mega=0)."""
import os, sys, runpy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import plangen_b200.engine as E
_orig = E.FastJanus.__init__
def patched(self, *a, **k):
    o = dict(k.get("options") or {}); o["use_mega"] = 0; k["options"] = o
    _orig(self, *a, **k)
E.FastJanus.__init__ = patched
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profile_step.py"), run_name="__main__")
