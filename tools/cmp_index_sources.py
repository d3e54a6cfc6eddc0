import json,sys
a=json.load(open('gpurun_out/ab_device.json')); b=json.load(open('gpurun_out/ab_files.json'))
ks=['num_aligned','sw_calls','sw_cells','windows','pos_entries','lis_calls','bucket_entries']
print({k:(a['counters'][k],b['counters'][k]) for k in ks})
print('SAME' if all(a['counters'][k]==b['counters'][k] for k in ks) else 'DIFFERENT')
for j in (a,b): print(j['index_source'], j['index_resident_s'], round(j['value']), round(j['e2e']['value']), j['kernel_ms_per_step'])
