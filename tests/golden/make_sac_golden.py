"""Extract the epoch-0 `trainer/*` known answers of the reference's committed 2020 runs into
tests/golden/sac_epoch0_known_answers.json.  Run in the build container: python tests/golden/make_sac_golden.py /root/reference
(the GPU box has no /root/reference; the JSON is what travels)."""
import csv, glob, json, os, sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
fams = ["Lift-Panda-OSC-POSE", "Door-Panda-JOINT-VELOCITY", "Stack-Sawyer-OSC-POSE", "TwoArmLift-PandaPanda-OSC-POSE"]
cols = ["trainer/Alpha", "trainer/Alpha Loss", "trainer/Log Pis Mean", "trainer/Log Pis Std", "trainer/Policy Loss", "trainer/QF1 Loss",
        "trainer/Q Targets Mean", "trainer/Q Targets Std", "trainer/Q1 Predictions Mean", "trainer/Policy mu Std", "trainer/Policy log std Mean",
        "replay_buffer/size", "exploration/num paths total", "evaluation/num paths total", "exploration/num steps total"]
out = {}
for fam in fams:
    for seed in (17, 59, 83, 129, 251):
        fs = glob.glob(os.path.join(ref, "runs", f"{fam}-SEED{seed}", "*", "progress.csv"))
        if not fs:
            continue
        rows = list(csv.DictReader(open(fs[0])))
        var = json.load(open(os.path.join(os.path.dirname(fs[0]), "variant.json")))
        out[f"{fam}-SEED{seed}"] = {"epoch0": {c: float(rows[0][c]) for c in cols}, "epoch1": {c: float(rows[1][c]) for c in cols[:2]},
                                    "trainer_kwargs": var["trainer_kwargs"], "batch_size": var["algorithm_kwargs"]["batch_size"]}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "sac_epoch0_known_answers.json"), "w"), indent=1, sort_keys=True)
print(len(out), "runs")
