import json,sys
d=json.loads(sys.stdin.read()); print(d["n_gpus"], "value", round(d["value"]), "ms/round", round(d["ms_per_step"],2)); print(d["rank0"]["host_seconds_per_phase"], d["rank0"]["seconds"])
