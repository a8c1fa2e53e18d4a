// stand-in for ncnn's datareader.h (see net.h in this directory)
