python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r2_u_config3.json 2> gpurun_out/r2_u_config3.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_u_config3.json").read().strip().splitlines()[-1])
print({k:round(v["utt_per_s"]) for k,v in d["token_min_logp_sweep"].items()}, {k:round(v["utt_per_s"]) for k,v in d["flat_logits_128_utterances"].items()})
PY
