mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -k "topk or dropin or autograd or layout" > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
tail -4 gpurun_out/pytest.log
python scripts/topk_probe.py
python scripts/topk_probe.py quick > gpurun_out/plain_topk.log 2>&1 && ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none -k regex:topk_cbsr --csv --log-file gpurun_out/topk_inst.csv python scripts/topk_probe.py quick > gpurun_out/ncu_topk2.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/topk_inst.csv")) if len(r)>10 and r[0].isdigit()]
names=["uniform","normal","relu_normal","quantised"]
i=0
for a,b in zip(rows[0::2], rows[1::2]):
    m={a[-3]:a[-1], b[-3]:b[-1]}
    print(names[i//3], [8,32,64][i%3], "inst/row", float(m["smsp__inst_executed.sum"].replace(",",""))/232965, "ns", m["gpu__time_duration.sum"])
    i+=1
PY
