"""build tuning variants of the library into variants/ (git-ignored): python profiles/scripts/build_variants.py name:FLAG,FLAG ..."""
import sys, os
sys.path.insert(0, ".")
from unsupervised_depth_opticalflow_egomotion_b200 import build
os.makedirs("variants", exist_ok=True)
for spec in sys.argv[1:]:
    name, flags = spec.split(":")
    out = os.path.abspath("variants/lib_%s.so" % name)
    build.build(force=True, extra_flags=["-D" + f for f in flags.split(",") if f], out=out)
    print("built", out)
