# follow-up of group_sweep.py: frames per group at the full sample counts of the shallow scenes, and the pool with it
export SWEEP_F="1 16 32 64 128 256"
timeout 100 python scripts/group_sweep.py c2:256 c3:256 c2:256:4194304 2>&1 | grep "^c[0-9]"
