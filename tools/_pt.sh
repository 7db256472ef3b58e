export E2E_AB_FULL=1
python - <<'P'
import os, subprocess, sys
sys.path.insert(0, "tools")
import e2e_ab
for t in (12, 4):
    e2e_ab.run(f"session/{t}", {"WF_HOST_THREADS": str(t)}, 1, 4096, 2000)
    e2e_ab.run(f"persistent/{t}", {"WF_HOST_THREADS": str(t)}, 2, 4096, 2000)
P
