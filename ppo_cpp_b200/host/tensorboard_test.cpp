// Writes a small event file with TensorboardWriter (host/tensorboard.hpp); tests/test_host_cpp.py reads it back with
// TensorBoard's own reader (which verifies both CRCs of every record).
#include "tensorboard.hpp"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    // CRC-32C known answers (RFC 3720 B.4): 32 zero bytes -> 0x8a9136aa, "123456789" -> 0xe3069283
    unsigned char zeros[32] = {0};
    if (TensorboardWriter::crc32c(zeros, 32) != 0x8a9136aau) return 3;
    if (TensorboardWriter::crc32c("123456789", 9) != 0xe3069283u) return 4;
    TensorboardWriter w(std::string(argv[1]) + "/", "PPO2");
    for (int i = 0; i < 150; ++i) w.write_scalar(1000.0 + i, i * 20, "episode_reward", 150.f / (i + 1));
    w.write_scalar(7, "other/scalar", -2.5f);
    std::cout << w.path() << std::endl;
    std::cout << "tensorboard_test OK" << std::endl;
    return 0;
}
