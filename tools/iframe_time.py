import os, sys
sys.path.insert(0, "/root/repo")
from libde265_b200 import synth
from libde265_b200.engine import Engine
p = synth.make_picture(3840, 2160, "I", seed=1000)
for ctas in (1, 2, 3):
    for ns in (256, 1024, 4096):
        os.environ["B200_INTRA_CTAS"] = str(ctas); os.environ["B200_POLL_NS"] = str(ns)
        eng = Engine(0); eng.enable_timing(True)
        ts = []
        for _ in range(3):
            eng.submit(p); eng.sync(); ts.append(eng.last_timing()["recon"])
        print("ctas/SM", ctas, "poll_ns", ns, "recon ms", ["%.2f" % t for t in ts], flush=True)
        eng.close()
