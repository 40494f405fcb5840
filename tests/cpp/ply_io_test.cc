// IOTest of the reference (test/teaser/io-test.cc:15-47) against the façade's PLYReader / PLYWriter, plus binary and
// big-endian / double-precision inputs.  usage: ply_io_test <cube.ply> <bunny.ply> <tmpdir>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "teaser/ply_io.h"

#define CHECK(c)                                                                  \
  do {                                                                            \
    if (!(c)) {                                                                   \
      std::cerr << "CHECK failed: " #c " at line " << __LINE__ << std::endl;     \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

static bool same(const teaser::PointCloud& a, const teaser::PointCloud& b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i)
    if (a[i].x != b[i].x || a[i].y != b[i].y || a[i].z != b[i].z) return false;
  return true;
}

int main(int argc, char** argv) {
  if (argc < 4) return 2;
  const std::string tmp = argv[3];
  teaser::PLYReader reader;
  teaser::PLYWriter writer;
  teaser::PointCloud cube;
  CHECK(reader.read(argv[1], cube) == 0);  // 8 vertices followed by a face element with list properties
  CHECK(cube.size() == 8);
  CHECK(cube[0].x == -1 && cube[0].y == -1 && cube[0].z == -1 && cube[6].x == 1 && cube[6].z == 1);
  // ImportPLY: write, read back, compare (ascii and binary)
  for (int binary = 0; binary < 2; ++binary) {
    const std::string out = tmp + (binary ? "/cube_bin.ply" : "/cube_ascii.ply");
    CHECK(writer.write(out, cube, binary != 0) == 0);
    teaser::PointCloud back;
    CHECK(reader.read(out, back) == 0);
    CHECK(same(cube, back));
  }
  teaser::PointCloud bunny;
  CHECK(reader.read(argv[2], bunny) == 0);  // ascii with extra float properties per vertex
  CHECK(bunny.size() == 1889);
  {
    const std::string out = tmp + "/bunny_bin.ply";
    CHECK(writer.write(out, bunny, true) == 0);
    teaser::PointCloud back;
    CHECK(reader.read(out, back) == 0);
    CHECK(same(bunny, back));
    const std::string out2 = tmp + "/bunny_ascii.ply";
    CHECK(writer.write(out2, bunny, false) == 0);
    teaser::PointCloud back2;
    CHECK(reader.read(out2, back2) == 0);
    CHECK(same(bunny, back2));  // 9 significant digits round-trip a float
  }
  {  // big-endian doubles with an int property in between
    const std::string out = tmp + "/be.ply";
    std::ofstream f(out, std::ios::binary);
    f << "ply\nformat binary_big_endian 1.0\nelement vertex 2\nproperty double x\nproperty int flag\nproperty double y\n"
         "property double z\nend_header\n";
    const double vals[2][3] = {{1.5, -2.25, 3.0}, {0.125, 1e-3, -7.0}};
    for (int r = 0; r < 2; ++r) {
      auto put_be = [&](const void* p, int n) {
        const unsigned char* b = static_cast<const unsigned char*>(p);
        for (int i = n - 1; i >= 0; --i) f.put(static_cast<char>(b[i]));
      };
      const int32_t flag = 7 + r;
      put_be(&vals[r][0], 8);
      put_be(&flag, 4);
      put_be(&vals[r][1], 8);
      put_be(&vals[r][2], 8);
    }
    f.close();
    teaser::PointCloud be;
    CHECK(reader.read(out, be) == 0);
    CHECK(be.size() == 2 && be[0].x == 1.5f && be[0].y == -2.25f && be[1].z == -7.0f && be[1].y == 1e-3f);
  }
  teaser::PointCloud none;
  CHECK(reader.read(tmp + "/does_not_exist.ply", none) == -1);  // "PLY reader returns -1"
  std::cout << "ok" << std::endl;
  return 0;
}
