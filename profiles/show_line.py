import sys,json
for line in sys.stdin:
    line=line.strip()
    if not line.startswith('{'): continue
    d=json.loads(line)
    ks={k['kernel']:round(k['ms_per_launch'],3) for k in d.get('roofline_kernels',[])}
    nb=d.get('notebook_mode',{})
    print(sys.argv[1], 'obj/s %.0f ms %.3f'%(d['value'],d['ms_per_step']), ks, '| notebook %.0f ms %.3f'%(nb.get('objects_per_s',0),nb.get('ms_per_step',0)), {k['kernel']:round(k['ms_per_launch'],3) for k in nb.get('kernels',[])}, 'e2e', d.get('e2e',{}).get('value'))
