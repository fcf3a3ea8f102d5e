#!/bin/bash
# dev: A/B of the warps-per-CTA / CTAs-per-SM geometries of the tuned grouping kernels (VGTKB_IG_* switches) on bench.py
out=gpurun_out/geo
mkdir -p $out
run() {  # name, env assignments...
    name=$1; shift
    env "$@" python bench.py --no-cpu-baseline --no-ref-gpu > $out/$name.json 2> $out/$name.err
}
run base0 VGTKB_IG_FWD1=0
for g in 63 44 102 53; do run fwd1_$g VGTKB_IG_FWD1=$g; done
for g in 62 53 43 44; do run fwd2_$g VGTKB_IG_FWD2=$g; done
run base1 VGTKB_IG_FWD1=0
for g in 63 44 102 53; do run bwd1_$g VGTKB_IG_BWD1=$g; done
for g in 63 53 44 102; do run bwd2_$g VGTKB_IG_BWD2=$g; done
run base2 VGTKB_IG_FWD1=0
python - <<'PY'
import json, glob, os
for f in sorted(glob.glob('gpurun_out/geo/*.json'), key=os.path.getmtime):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(os.path.basename(f), 'FAILED', e); continue
    kt = d['kernel_table']
    st = d['shape_table']
    fw = {k: v['ms_per_step'] for k, v in st.items() if k.startswith('inter_conv_forward')}
    bw = {k: v['ms_per_step'] for k, v in st.items() if k.startswith('inter_conv_backward')}
    f1 = sum(v for k, v in fw.items() if ', 16, 60' in k); f2 = sum(v for k, v in fw.items() if ', 32, 60' in k)
    b1 = sum(v for k, v in bw.items() if ', 16, 60' in k); b2 = sum(v for k, v in bw.items() if ', 32, 60' in k)
    print('%-10s ms/step %.3f  fwd nn16 %.3f nn32 %.3f | bwd nn16 %.3f nn32 %.3f | clk %s' % (
        os.path.basename(f)[:-5], d['ms_per_step'], f1, f2, b1, b2, d['clocks'].get('sm_mhz')))
PY
