#!/bin/bash
# profile evidence: launch lists (inference step, train step, cfg4) and `ncu --set full` captures of one launch of every kernel family.
# gpurun copies back at most 64 MiB: summaries are produced ON the box, raw reports are kept only while they fit.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
O=gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_fwd.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-train-block > $O/r02_prof_a.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2200 --csv --log-file $O/r02_launches_train.csv \
  python bench.py --train --n-sources 3 --batch 8 --steps 1 --warmup 2 > $O/r02_prof_b.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r02_launches_cfg4.csv \
  python bench.py --config cfg4 --steps 1 --warmup 3 --no-cpu-baseline > $O/r02_prof_c.log 2>&1
# (1) the three TMA-fed kernels with source: launches of the 3rd forward: pw1 (block 5), pw2 (block 5, d=32), ..., mask+decoder
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_pw_tma --launch-skip 108 --launch-count 2 -o $O/r02_full_pw \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-block > $O/r02_prof_d.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:k_pw_tma --launch-skip 145 --launch-count 1 -o $O/r02_full_maskdec \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-block > $O/r02_prof_e.log 2>&1
# (2) every other kernel of the forward + loss, one launch each (third step)
timeout 900 ncu --set full --clock-control none -k regex:"k_encoder|k_pw_umma|k_skip_reduce|k_pit_|k_fold_batch|k_build_wimg_batch|k_scale_|k_batch_mean" \
  --launch-skip 36 --launch-count 14 -o $O/r02_full_rest python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-train-block > $O/r02_prof_f.log 2>&1
for r in r02_full_pw r02_full_maskdec r02_full_rest; do
  [ -f $O/$r.ncu-rep ] && python tools/ncu_summary.py $O/$r.ncu-rep "$r" > $O/$r.md 2>&1
done
[ -f $O/r02_full_pw.ncu-rep ] && { python tools/ncu_roles.py $O/r02_full_pw.ncu-rep 0 > $O/r02_roles_pw1.txt 2>&1; python tools/ncu_roles.py $O/r02_full_pw.ncu-rep 1 > $O/r02_roles_pw2.txt 2>&1; }
du -sm $O | tail -1
# keep the copy-back under the 64 MiB limit
sz=$(du -sm $O | cut -f1); if [ "$sz" -gt 55 ]; then rm -f $O/r02_full_rest.ncu-rep; fi
sz=$(du -sm $O | cut -f1); if [ "$sz" -gt 55 ]; then rm -f $O/r02_full_maskdec.ncu-rep; fi
sz=$(du -sm $O | cut -f1); if [ "$sz" -gt 55 ]; then rm -f $O/r02_full_pw.ncu-rep; fi
ls -la $O | grep "r02_"; cat $O/r02_full_pw.md $O/r02_full_maskdec.md $O/r02_full_rest.md 2>/dev/null | head -60
