import sys
exec(open("/root/repo/tools/conv_microbench.py").read().split("for stats in")[0].replace('sys.path.insert(0, ".")', 'sys.path.insert(0, "/root/repo")'))
bench("stem rowk 8->64 @256 B16", 16, 8, 64, 256, 7, stats=False, halo=False, rowk=True)
bench("skipper 64+64->64 @256 B16", 16, 64, 64, 256, 3, stats=False, halo=False, cin1=64)
bench("skipper 128+128->128 @128 B16", 16, 128, 128, 128, 3, stats=False, halo=False, cin1=128)
bench("res 512->512 @32 B16", 16, 512, 512, 32, 3, stats=False, halo=False)
