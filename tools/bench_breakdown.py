import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
b=j['breakdown']
print(sys.argv[1], round(j['value']), 'clk', j['clocks']['sm_mhz'], ' '.join(f"{k}={b[k]['ms_per_step']:.1f}" for k in ('dec2.ru.conv1','dec3.ru.conv1','dec1.ru.conv1','dec2.ru.conv7','dec3.ru.conv7','dec2.convt','dec3.convt')))
