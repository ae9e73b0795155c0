"""`ptxas -v` table from the build logs the Makefile leaves next to the objects (avatarcap_b200/csrc/*.o.log):
    python profiles/ptxas_summary.py > profiles/r1_ptxas_v.md"""
import glob, os, re, subprocess
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
print('# ptxas -v (sm_100a, nvcc 12.9) per kernel: registers, spills, static shared memory, stack\n')
print('| file | kernel | registers | spill st/ld (B) | static smem (B) | stack (B) | barriers |\n|---|---|---|---|---|---|---|')
for log in sorted(glob.glob(os.path.join(root, 'avatarcap_b200', 'csrc', '*.o.log'))):
    txt = open(log).read()
    for m in re.finditer(r"Compiling entry function '([^']+)' for 'sm_100a'.*?(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\s*\n"
                         r"ptxas info\s*: Used (\d+) registers, used (\d+) barriers(?:, (\d+) bytes cumulative stack size)?(?:, (\d+) bytes smem)?", txt, re.S):
        mangled, stack, sst, sld, regs, bars, cstack, smem = m.groups()
        try:
            name = subprocess.run(['c++filt', mangled], capture_output=True, text=True).stdout.strip()
        except Exception:
            name = mangled
        name = re.sub(r'\(anonymous namespace\)::', '', name).split('(')[0].replace('void ', '')
        print('| %s | `%s` | %s | %s / %s | %s | %s | %s |' % (os.path.basename(log)[:-6] + '.cu', name, regs, sst, sld, smem or 0, cstack or stack, bars))
