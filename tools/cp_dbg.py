import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
cpm = importlib.import_module("qwen3-tts-axera-russian_b200.code_predictor")
cfg = cpm.CPConfig()
w = cpm.init_weights(cfg, 0)
rng = np.random.default_rng(0)
hs = rng.standard_normal((16, cfg.hidden)).astype(np.float32)
es = rng.standard_normal((16, cfg.hidden)).astype(np.float32)
for seq in (["graph", "batch2"], ["graph", "persistent", "graph", "batch2"], ["graph", "level1", "batch2"], ["graph", "batch4", "batch8", "level1"]):
    cp = cpm.CodePredictor(cfg, w)
    try:
        for op in seq:
            if op in ("graph", "persistent"):
                cp.set_option("predict", op)
                cp.predict(hs[0], es[0], 0.1, 50, seed=1)
            elif op == "level1":
                cp.reset(); cp.step(hs[:1], 0); cp.logits(0)
            else:
                B = int(op[5:])
                cp.predict_batch(hs[:B], es[:B], 0.1, 50)
        print(seq, "OK")
    except Exception as e:
        print(seq, "FAILED at", op, e)
    cp.close()
