// Harness for tests/test_image_ingest.py: exposes the host layer's PNG decoder and Lanczos3 resampler (arendur_b200/csrc/host/image_io.hpp).
//   test_image_io decode FILE            -> "w h ch" then the pixel bytes as hex
//   test_image_io resize FILE NW NH      -> the same for resize_lanczos3(decode(FILE), NW, NH)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "../../arendur_b200/csrc/host/image_io.hpp"

int main(int argc, char** argv) {
    if (argc < 3) return 2;
    arnhost::Image8 img; std::string err;
    if (!arnhost::image_decode(argv[2], &img, &err)) { std::printf("ERROR %s\n", err.c_str()); return 1; }
    if (!std::strcmp(argv[1], "resize") && argc >= 5) img = arnhost::resize_lanczos3(img, (uint32_t)std::atoi(argv[3]), (uint32_t)std::atoi(argv[4]));
    std::printf("%u %u %u\n", img.w, img.h, img.ch);
    for (uint8_t b : img.px) std::printf("%02x", b);
    std::printf("\n");
    return 0;
}
