// Declaration-only stand-in for ncnn's net.h (SkySegment/include/SkyRegionDetect.h:19): the sky-mask filter kernel's
// translation unit includes the header of the segmentation class, which holds an ncnn::Net by value. ncnn is not in this
// image and nothing of it is called. This file is ours (test infrastructure), not reference code.
#ifndef MPMVS_REF_SHIM_NCNN_NET_H
#define MPMVS_REF_SHIM_NCNN_NET_H
namespace ncnn {
struct Option { bool use_vulkan_compute; };
class Net {
  public:
    Option opt;
};
}  // namespace ncnn
#endif
