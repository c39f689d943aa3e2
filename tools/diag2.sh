B200_KLU_TRACE=1 timeout 300 python - <<'PY' > gpurun_out/r2e_diag2.log 2>&1
import sys; sys.path.insert(0,'.')
from oracle.pyoracle import run_ref_klu
from tests.decks import variants
from xictools_b200 import engine
r = run_ref_klu(variants()["rlc_series"].to_spice(), engine.LIB_PATH)
print(r["stderr"][:6000])
PY
head -120 gpurun_out/r2e_diag2.log
