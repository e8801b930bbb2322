"""Prints the key metrics of every kernel in an .ncu-rep (reads `ncu --page details --csv`)."""
import csv, subprocess, sys
WANT = ['Duration', 'Executed Ipc Active', 'Registers Per Thread', 'Achieved Occupancy', 'L1/TEX Hit Rate', 'L2 Hit Rate',
        'Executed Instructions', 'Avg. Active Threads Per Warp', 'Issue Slots Busy', 'DRAM Throughput', 'Theoretical Occupancy',
        'Warp Cycles Per Issued Instruction', 'No Eligible', 'Compute (SM) Throughput', 'Memory Throughput',
        'Mem Busy', 'Max Bandwidth', 'L1/TEX Cache Throughput', 'L2 Cache Throughput', 'Dynamic Shared Memory Per Block']
out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'details', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
last = None
for r in rows[1:]:
    d = dict(zip(h, r))
    key = (d['ID'], d['Kernel Name'][:60])
    if key != last:
        print('---', *key)
        last = key
    if d['Metric Name'] in WANT:
        print('   %-40s %s %s' % (d['Metric Name'], d['Metric Value'], d['Metric Unit']))
