mkdir -p gpurun_out
R="python scripts/run_case.py"
{
LDPC_PLAN_EFFORT=16 $R --algo MSA
LDPC_PLAN_EFFORT=16 $R --algo SPA --cw 0
$R --code 1200_rho_x5_rand_ldpc_1 --channel bsc --snr 0.06 --algo SPA --cw 0 --max-iter 100
$R --code 1200_rho_x5_rand_ldpc_1 --channel bsc --snr 0.06 --algo SPA --cw 0 --max-iter 100 --streaming
$R --code 1200_rho_x5_rand_ldpc_1 --channel bsc --snr 0.06 --algo SPA --cw 0 --max-iter 10
$R --code 1200_rho_x5_rand_ldpc_1 --channel bsc --snr 0.04 --algo SPA --cw 0 --max-iter 100
} 2>&1 | tee gpurun_out/cases.txt
