# sweeps of the bucket reduction's level passes for batched commits (tools/acc_waves_sweep.py --child)
for v in 3 4 5; do echo "== MIRA_RED_LOG_M=$v"; MIRA_RED_LOG_M=$v MIRA_ACC_WAVES= python tools/acc_waves_sweep.py --child 2>&1 | tail -1; done
