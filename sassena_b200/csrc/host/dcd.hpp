// dcd.hpp — CHARMM DCD trajectory reader / writer feeding the stager.
// Reader: DCDFrameset (reference src/sample/frames.cpp:272-436; header struct include/sample/frames.hpp:152-164;
// first/last/stride trimming FileFrameset::trim_index frames.cpp:224-245).
// Writer: DCDCoordinateWriter layout (reference src/stager/coordinate_writer.cpp:37-144), used for stager.dump and to
// produce synthetic trajectories the reference itself can read.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace sassena {

class DCDFrameset {
    std::string filename_;
    FILE *f_ = nullptr;
    std::vector<int64_t> frameset_index_;  // byte offset of every (kept) frame
    int64_t init_byte_pos = 0, block_size_byte = 0;
    int64_t block1_byte_offset = 0, x_byte_offset = 0, y_byte_offset = 0, z_byte_offset = 0, block2_byte_offset = 0;
    int32_t flag_ext_block1 = 0, flag_ext_block2 = 0;
    std::vector<float> buf_;
    void parse_header(const std::string &fn);

   public:
    size_t number_of_frames = 0, number_of_atoms = 0;
    explicit DCDFrameset(const std::string &fn);
    ~DCDFrameset();
    DCDFrameset(const DCDFrameset &) = delete;
    DCDFrameset &operator=(const DCDFrameset &) = delete;
    static bool detect(const std::string &fn);
    // keep frames i with i >= first, (i <= last if last_set), i % stride == 0   (frames.cpp:224-245)
    void trim_index(size_t first, size_t last, bool last_set, size_t stride);
    // xyz: [number_of_atoms][3] interleaved (the stager's layout); unitcell (6 doubles) may be NULL
    void read_frame(size_t framenumber, float *xyz, double *unitcell = nullptr);
    // frames [first, first+count) of the trimmed index into out[count][NA][3]
    void read_frames(size_t first, size_t count, float *out);
    bool has_unitcell() const { return flag_ext_block1 != 0; }
};

class DCDCoordinateWriter {
    std::string file_;
    size_t blocks_, entries_;
    int64_t data_offset_ = 0;

   public:
    DCDCoordinateWriter(const std::string &file, size_t blocks, size_t entries) : file_(file), blocks_(blocks), entries_(entries) {}
    void init();     // header (ext block 1 on, empty title)
    void prepare();  // data offset = end of header
    // data: [myblocks][entries][3] interleaved; frames blockoffset .. blockoffset+myblocks
    void write(const float *data, size_t blockoffset, size_t myblocks);
};

}  // namespace sassena
