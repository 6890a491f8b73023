"""TEST INFRASTRUCTURE ONLY -- golden learning-rate sequences from the reference's own scheduler classes.

Run in the build container (``python oracle/make_golden_schedulers.py``): loads
``/root/reference/biapy/engine/schedulers/warmup_cosine_decay.py`` and ``warmup_reduce_on_plateau.py`` unmodified, drives them
the way ``train_one_epoch`` does (``epoch + step / steps_per_epoch``, ``train_engine.py:113-116``) and writes
``tests/golden/schedulers.json``."""
import importlib.util
import json
import os

REF = "/root/reference/biapy/engine/schedulers"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "schedulers.json")


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Opt:
    def __init__(self):
        self.param_groups = [{"lr": 0.0}, {"lr": 0.0, "lr_scale": 0.5}]


def main():
    wc = load("warmup_cosine_decay").WarmUpCosineDecayScheduler
    wr = load("warmup_reduce_on_plateau").WarmUpReduceOnPlateauScheduler
    out = {"warmupcosine": [], "warmupreduceonplateau": []}
    for lr, min_lr, warm, epochs, spe in [(1e-3, 1e-5, 5, 40, 7), (2e-4, 0.0, 1, 10, 3), (1e-2, 1e-4, 10, 360, 2)]:
        s, o = wc(lr=lr, min_lr=min_lr, warmup_epochs=warm, epochs=epochs), Opt()
        seq = []
        for e in range(epochs):
            for st in range(spe):
                r = s.adjust_learning_rate(o, st / spe + e)
                seq.append([r, o.param_groups[0]["lr"], o.param_groups[1]["lr"]])
        out["warmupcosine"].append({"lr": lr, "min_lr": min_lr, "warmup_epochs": warm, "epochs": epochs, "steps_per_epoch": spe, "seq": seq})
    for lr, epochs in [(1e-3, 8), (1e-3, 60), (5e-4, 150), (1e-4, 360)]:
        s, o = wr(lr=lr, epochs=epochs), Opt()
        seq = []
        for e in range(epochs + 3):
            r = s.adjust_learning_rate(o, e + 0.5)
            seq.append([r, o.param_groups[0]["lr"], o.param_groups[1]["lr"]])
        out["warmupreduceonplateau"].append({"lr": lr, "epochs": epochs, "table": [float(v) for v in s.LR], "seq": seq})
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
