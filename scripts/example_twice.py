import importlib.util, os, sys
spec = importlib.util.spec_from_file_location("omr_example", os.path.join(os.getcwd(), "examples", "omr.py"))
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
mod.main(["-p", "65536"]); print("---- second run, same process"); mod.main(["-p", "65536", "--no-warm-up"])
