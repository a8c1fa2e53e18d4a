// stand-in for ncnn's benchmark.h (see net.h in this directory)
