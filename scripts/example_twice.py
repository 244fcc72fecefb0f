"""examples/omr.py --payload-count 65536 twice in one process: first run (after the example's own D = 8 warm-up pass) and a fully warm
second run — profiles/r2_example_omr_65536.txt."""
import importlib.util, os, sys
spec = importlib.util.spec_from_file_location("omr_example", os.path.join(os.getcwd(), "examples", "omr.py"))
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
mod.main(["-p", "65536"]); print("---- second run, same process"); mod.main(["-p", "65536", "--no-warm-up"])
