for mode in weak strong strong-b; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 --scaling $mode $( [ $mode != weak ] && echo --no-e2e --no-extras ) --no-cpu > gpurun_out/r2d_n8_$mode.json 2> gpurun_out/r2d_n8_$mode.err
tail -c 200 gpurun_out/r2d_n8_$mode.err
python -c "
import json; d=json.loads(open('gpurun_out/r2d_n8_$mode.json').read().strip().splitlines()[-1]); print('$mode', d['value'], d['ms_per_step'], d.get('parity',{}).get('ok'), d.get('phases',{}).get('ms_max_over_ranks'), d.get('e2e',{}).get('value'), d.get('e2e',{}).get('ms_per_step'), d['roofline']['frac'])"
done
