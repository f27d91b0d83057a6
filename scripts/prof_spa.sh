# ncu capture of the on-chip kernel running float32 sum-product (run under gpurun)
mkdir -p gpurun_out
python scripts/run_case.py --algo SPA --cw 0 | tee gpurun_out/spa_case.txt
python scripts/run_case.py --algo SPA --cw 0 --streaming | tee -a gpurun_out/spa_case.txt
python scripts/run_case.py --algo SPA --cw 0 --snr 3.0 | tee -a gpurun_out/spa_case.txt
ncu --set full --clock-control none --import-source on -k regex:resident_bp -s 2 -c 1 -o gpurun_out/resident_spa_r1h -f python scripts/run_case.py --algo SPA --cw 0 --steps 1 > /dev/null 2>&1
ncu -i gpurun_out/resident_spa_r1h.ncu-rep --page raw --csv > gpurun_out/resident_spa_r1h_raw.csv
ls -la gpurun_out
