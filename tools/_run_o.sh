python tools/beam_perf.py --utts 8192 --iters 7 > gpurun_out/r2_o_perf.log 2>&1
python tools/beam_perf.py --utts 1776 --kind flat --iters 3 >> gpurun_out/r2_o_perf.log 2>&1
python tools/beam_perf.py --utts 2048 --beam 256 --iters 3 >> gpurun_out/r2_o_perf.log 2>&1
python tools/beam_perf.py --utts 2048 --beam 512 --iters 3 >> gpurun_out/r2_o_perf.log 2>&1
python bench.py --config 3 --steps 5 --warmup 3 > gpurun_out/r2_o_config3.json 2> gpurun_out/r2_o_config3.err
grep utts gpurun_out/r2_o_perf.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_o_config3.json").read().strip().splitlines()[-1])
print({k:round(v["utt_per_s"]) for k,v in d["token_min_logp_sweep"].items()}, {k:round(v["utt_per_s"]) for k,v in d["flat_logits_128_utterances"].items()})
PY
