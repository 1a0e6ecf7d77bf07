// Host-side VTU writer of the output side: the uncompressed branch of the reference's save_vtu
// (src/mesh/io.jl:167-276) with the XML layout of src/tools/xml.jl:253-316, fed from flat arrays instead of the
// node / element object graph.  DataArray rows are formatted exactly like get_array_node! (io.jl:150-163):
// floats "%20.10e" of Float32(value), integers followed by two blanks, rows separated by newlines, and every line's
// leading blanks replaced by the 3-blank-per-level indentation (xml.jl:280).  Rows are formatted by all host threads.
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "amaru_internal.h"

namespace {

const char *TAB = "   ";

std::string indent(int level) {
    std::string s;
    for (int i = 0; i < level; i++) s += TAB;
    return s;
}

int vtk_type(int shape) {   // src/shape/shape.jl:66-96
    switch (shape) {
    case AMARU_SHAPE_QUAD4: return 9;
    case AMARU_SHAPE_QUAD8: return 23;
    case AMARU_SHAPE_HEX8: return 12;
    case AMARU_SHAPE_HEX20: return 25;
    case AMARU_SHAPE_TET10: return 24;
    }
    return -1;
}

// type: 0 Float64, 1 Int64, 2 Int32, 3 UInt64
template <class T>
void format_rows(const T *a, int64_t lo, int64_t hi, int nc, bool isfloat, const std::string &ind, std::string &out) {
    char buf[64];
    out.clear();
    out.reserve((size_t)(hi - lo) * (size_t)(nc * 21 + ind.size() + 1));
    for (int64_t i = lo; i < hi; i++) {
        out += ind;
        for (int j = 0; j < nc; j++) {
            if (isfloat) {
                const int n = std::snprintf(buf, sizeof buf, "%20.10e", (double)(float)a[i * nc + j]);
                const char *p = buf;
                int len = n;
                if (j == 0)
                    while (*p == ' ') { p++; len--; }   // leading blanks of a line are replaced by the indentation
                out.append(p, (size_t)len);
            } else {
                const int n = std::is_unsigned<T>::value ? std::snprintf(buf, sizeof buf, "%llu  ", (unsigned long long)a[i * nc + j])
                                                         : std::snprintf(buf, sizeof buf, "%lld  ", (long long)a[i * nc + j]);
                out.append(buf, (size_t)n);
            }
        }
        out += '\n';
    }
}

template <class T>
void write_array(FILE *f, const char *type, const char *name, const T *a, int64_t nrows, int nc, bool isfloat, int level) {
    const std::string ind = indent(level);
    std::fprintf(f, "%s<DataArray type=\"%s\" Name=\"%s\" NumberOfComponents=\"%d\" format=\"ascii\"", ind.c_str(), type, name, nc);
    if (nrows == 0) {
        std::fprintf(f, "/>\n");
        return;
    }
    std::fprintf(f, ">\n");
    const std::string cind = indent(level + 1);
    int nt = amaru_host_threads();
    if (nrows * nc < 65536) nt = 1;
    std::vector<std::string> parts((size_t)nt);
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) {
        const int64_t lo = nrows * t / nt, hi = nrows * (t + 1) / nt;
        if (nt == 1) format_rows(a, lo, hi, nc, isfloat, cind, parts[0]);
        else th.emplace_back([&, t, lo, hi] { format_rows(a, lo, hi, nc, isfloat, cind, parts[(size_t)t]); });
    }
    for (auto &x : th) x.join();
    for (auto &p : parts) std::fwrite(p.data(), 1, p.size(), f);
    std::fprintf(f, "%s</DataArray>\n", ind.c_str());
}

void write_any(FILE *f, const char *name, int type, int nc, const void *data, int64_t nrows, int level) {
    if (type == 0) write_array(f, "Float64", name, static_cast<const double *>(data), nrows, nc, true, level);
    else if (type == 1) write_array(f, "Int64", name, static_cast<const int64_t *>(data), nrows, nc, false, level);
    else if (type == 3) write_array(f, "UInt64", name, static_cast<const uint64_t *>(data), nrows, nc, false, level);
    else write_array(f, "Int32", name, static_cast<const int32_t *>(data), nrows, nc, false, level);
}

}  // namespace

extern "C" int amaru_write_vtu(const char *filename, const char *desc, int64_t nnodes, const double *coords, int nbatches,
                               const int32_t *batch_shape, const int64_t *batch_nelem, const int32_t *conn,
                               int npoint_arrays, const char *const *point_names, const int32_t *point_type,
                               const int32_t *point_ncomp, const void *const *point_data, int ncell_arrays,
                               const char *const *cell_names, const int32_t *cell_type, const int32_t *cell_ncomp,
                               const void *const *cell_data, char *msg, int msglen) {
    auto fail = [&](int code, const std::string &s) {
        if (msg && msglen > 0) std::snprintf(msg, (size_t)msglen, "%s", s.c_str());
        return code;
    };
    if (msg && msglen > 0) msg[0] = 0;
    if (!filename || !coords || !batch_shape || !batch_nelem || !conn || nnodes <= 0 || nbatches <= 0)
        return fail(AMARU_ERR_ARG, "amaru_write_vtu: null / empty argument");
    for (int i = 0; i < npoint_arrays; i++)
        if (!point_names || !point_type || !point_ncomp || !point_data || !point_names[i] || !point_data[i] || point_ncomp[i] < 1 ||
            point_type[i] < 0 || point_type[i] > 3)
            return fail(AMARU_ERR_ARG, "amaru_write_vtu: bad point array");
    for (int i = 0; i < ncell_arrays; i++)
        if (!cell_names || !cell_type || !cell_ncomp || !cell_data || !cell_names[i] || !cell_data[i] || cell_ncomp[i] < 1 ||
            cell_type[i] < 0 || cell_type[i] > 3)
            return fail(AMARU_ERR_ARG, "amaru_write_vtu: bad cell array");
    int64_t ncells = 0, nconn = 0;
    std::vector<int> nn((size_t)nbatches);
    for (int b = 0; b < nbatches; b++) {
        ShapeInfo si;
        if (vtk_type(batch_shape[b]) < 0 || !amaru_shape_info(batch_shape[b], si))
            return fail(AMARU_ERR_UNSUPPORTED, "amaru_write_vtu: cell shape outside the hot path");
        nn[(size_t)b] = si.nn;
        ncells += batch_nelem[b];
        nconn += batch_nelem[b] * si.nn;
    }
    FILE *f = std::fopen(filename, "w");
    if (!f) return fail(AMARU_ERR_ARG, std::string("amaru_write_vtu: cannot open ") + filename);
    std::vector<char> iobuf(1 << 22);
    std::setvbuf(f, iobuf.data(), _IOFBF, iobuf.size());
    std::fprintf(f, "<?xml version=\"1.0\" encoding=\"UTF-8\"?>\n");
    std::fprintf(f, "<!-- %s -->\n", desc ? desc : "");
    std::fprintf(f, "<VTKFile type=\"UnstructuredGrid\" version=\"1.0\" byte_order=\"LittleEndian\" header_type=\"UInt64\" "
                    "compressor=\"vtkZLibDataCompressor\">\n");
    std::fprintf(f, "%s<UnstructuredGrid>\n", indent(1).c_str());
    std::fprintf(f, "%s<Piece NumberOfPoints=\"%lld\" NumberOfCells=\"%lld\">\n", indent(2).c_str(), (long long)nnodes, (long long)ncells);
    std::fprintf(f, "%s<Points>\n", indent(3).c_str());
    write_array(f, "Float64", "Points", coords, nnodes, 3, true, 4);
    std::fprintf(f, "%s</Points>\n", indent(3).c_str());
    std::fprintf(f, "%s<Cells>\n", indent(3).c_str());
    write_array(f, "Int32", "connectivity", conn, nconn, 1, false, 4);
    {
        std::vector<int32_t> offsets((size_t)ncells), types((size_t)ncells);
        int64_t c = 0;
        int32_t off = 0;
        for (int b = 0; b < nbatches; b++)
            for (int64_t e = 0; e < batch_nelem[b]; e++, c++) {
                off += nn[(size_t)b];
                offsets[(size_t)c] = off;
                types[(size_t)c] = vtk_type(batch_shape[b]);
            }
        write_array(f, "Int32", "offsets", offsets.data(), ncells, 1, false, 4);
        write_array(f, "Int32", "types", types.data(), ncells, 1, false, 4);
    }
    std::fprintf(f, "%s</Cells>\n", indent(3).c_str());
    if (npoint_arrays > 0) {
        std::fprintf(f, "%s<PointData>\n", indent(3).c_str());
        for (int i = 0; i < npoint_arrays; i++) write_any(f, point_names[i], point_type[i], point_ncomp[i], point_data[i], nnodes, 4);
        std::fprintf(f, "%s</PointData>\n", indent(3).c_str());
    }
    if (ncell_arrays > 0) {
        std::fprintf(f, "%s<CellData>\n", indent(3).c_str());
        for (int i = 0; i < ncell_arrays; i++) write_any(f, cell_names[i], cell_type[i], cell_ncomp[i], cell_data[i], ncells, 4);
        std::fprintf(f, "%s</CellData>\n", indent(3).c_str());
    }
    std::fprintf(f, "%s</Piece>\n", indent(2).c_str());
    std::fprintf(f, "%s</UnstructuredGrid>\n", indent(1).c_str());
    std::fprintf(f, "</VTKFile>\n");
    const bool bad = std::ferror(f) != 0;
    if (std::fclose(f) != 0 || bad) return fail(AMARU_ERR_ARG, std::string("amaru_write_vtu: write error on ") + filename);
    return AMARU_OK;
}
