// PCD v0.7 reader for the frames that feed the path (host code, no CUDA): what pcl::io::loadPCDFile into a
// pcl::PointCloud<pcl::PointXYZI> followed by Dataloader::convert gives the reference
// (reference src/dataloader.cpp:87-126, 139): records on the PointCloud2 wire are sizeof(PointXYZI) = 32
// bytes apart with x at 0, y at 4, z at 8, intensity at 16 (conversions.cpp:62-85 reads them back).
// Supported: DATA binary and DATA ascii, fields x y z (+ optional intensity) of TYPE F SIZE 4 COUNT 1, any
// other fields are skipped by their SIZE*COUNT. DATA binary_compressed is rejected (not used by data/*.pcd).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace lb
{

struct PcdHeader
{
    uint64_t points{0};
    uint32_t record_bytes{0};               // bytes of one binary record
    int off_x{-1}, off_y{-1}, off_z{-1}, off_i{-1}; // byte offsets (binary) inside a record
    int col_x{-1}, col_y{-1}, col_z{-1}, col_i{-1}; // token columns (ascii)
    uint32_t columns{0};
    int data_kind{-1}; // 0 ascii, 1 binary
    long data_pos{0};
};

// returns an empty string on success, else the reason
inline std::string pcd_parse_header(FILE *f, PcdHeader *h)
{
    std::vector<std::string> fields, types;
    std::vector<uint32_t> sizes, counts;
    uint64_t width = 0, height = 1;
    bool have_points = false;
    char line[4096];
    auto split = [](const char *s) {
        std::vector<std::string> out;
        std::string cur;
        for (; *s; ++s)
        {
            if (*s == ' ' || *s == '\t' || *s == '\r' || *s == '\n')
            {
                if (!cur.empty())
                    out.push_back(cur);
                cur.clear();
            }
            else
                cur.push_back(*s);
        }
        if (!cur.empty())
            out.push_back(cur);
        return out;
    };
    while (std::fgets(line, sizeof(line), f))
    {
        if (line[0] == '#')
            continue;
        const std::vector<std::string> tok = split(line);
        if (tok.empty())
            continue;
        const std::string &key = tok[0];
        if (key == "FIELDS" || key == "COLUMNS")
            fields.assign(tok.begin() + 1, tok.end());
        else if (key == "SIZE")
            for (size_t i = 1; i < tok.size(); ++i)
                sizes.push_back(static_cast<uint32_t>(std::strtoul(tok[i].c_str(), nullptr, 10)));
        else if (key == "TYPE")
            types.assign(tok.begin() + 1, tok.end());
        else if (key == "COUNT")
            for (size_t i = 1; i < tok.size(); ++i)
                counts.push_back(static_cast<uint32_t>(std::strtoul(tok[i].c_str(), nullptr, 10)));
        else if (key == "WIDTH" && tok.size() > 1)
            width = std::strtoull(tok[1].c_str(), nullptr, 10);
        else if (key == "HEIGHT" && tok.size() > 1)
            height = std::strtoull(tok[1].c_str(), nullptr, 10);
        else if (key == "POINTS" && tok.size() > 1)
        {
            h->points = std::strtoull(tok[1].c_str(), nullptr, 10);
            have_points = true;
        }
        else if (key == "DATA" && tok.size() > 1)
        {
            if (tok[1] == "ascii")
                h->data_kind = 0;
            else if (tok[1] == "binary")
                h->data_kind = 1;
            else
                return "unsupported DATA " + tok[1];
            h->data_pos = std::ftell(f);
            break;
        }
    }
    if (h->data_kind < 0)
        return "no DATA line";
    if (!have_points)
        h->points = width * height;
    if (fields.empty() || sizes.size() != fields.size() || types.size() != fields.size())
        return "FIELDS / SIZE / TYPE missing or inconsistent";
    if (counts.empty())
        counts.assign(fields.size(), 1u);
    if (counts.size() != fields.size())
        return "COUNT inconsistent with FIELDS";
    uint64_t off = 0, col = 0; // 64-bit: SIZE and COUNT come from the file
    constexpr uint64_t kMaxRecordBytes = 65536u;
    for (size_t i = 0; i < fields.size(); ++i)
    {
        if (sizes[i] == 0u || sizes[i] > 8u || counts[i] == 0u || counts[i] > kMaxRecordBytes)
            return "SIZE / COUNT out of range";
        const bool f32 = types[i] == "F" && sizes[i] == 4u && counts[i] == 1u;
        int *o = nullptr, *c = nullptr;
        if (fields[i] == "x")
            o = &h->off_x, c = &h->col_x;
        else if (fields[i] == "y")
            o = &h->off_y, c = &h->col_y;
        else if (fields[i] == "z")
            o = &h->off_z, c = &h->col_z;
        else if (fields[i] == "intensity")
            o = &h->off_i, c = &h->col_i;
        if (o)
        {
            if (!f32)
                return "field " + fields[i] + " is not a single float32";
            *o = static_cast<int>(off);
            *c = static_cast<int>(col);
        }
        off += static_cast<uint64_t>(sizes[i]) * counts[i];
        col += counts[i];
        if (off > kMaxRecordBytes)
            return "point record larger than 64 KB";
    }
    h->record_bytes = static_cast<uint32_t>(off);
    h->columns = static_cast<uint32_t>(col);
    if (h->off_x < 0 || h->off_y < 0 || h->off_z < 0)
        return "fields x, y, z are required";
    return std::string();
}

// writes `stride_bytes`-spaced records: 16 = packed (x, y, z, intensity); 32 = pcl::PointXYZI wire layout
// (x, y, z, 1.0f, intensity, 0, 0, 0)
inline void pcd_store(uint8_t *dst, uint32_t stride_bytes, float x, float y, float z, float intensity)
{
    if (stride_bytes == 16u)
    {
        const float rec[4] = {x, y, z, intensity};
        std::memcpy(dst, rec, 16);
    }
    else
    {
        const float rec[8] = {x, y, z, 1.0f, intensity, 0.0f, 0.0f, 0.0f};
        std::memcpy(dst, rec, 32);
    }
}

inline std::string pcd_read(const char *path, void *points_out, uint64_t capacity_points, uint32_t stride_bytes,
                            uint64_t *n_points_out)
{
    if (!path || !n_points_out || (stride_bytes != 16u && stride_bytes != 32u))
        return "bad argument";
    FILE *f = std::fopen(path, "rb");
    if (!f)
        return std::string("cannot open ") + path;
    PcdHeader h;
    std::string err = pcd_parse_header(f, &h);
    if (!err.empty())
    {
        std::fclose(f);
        return err;
    }
    *n_points_out = h.points;
    if (!points_out) // size query
    {
        std::fclose(f);
        return std::string();
    }
    if (h.points > capacity_points)
    {
        std::fclose(f);
        return "capacity too small";
    }
    uint8_t *dst = static_cast<uint8_t *>(points_out);
    if (h.data_kind == 1)
    {
        std::vector<uint8_t> buf(static_cast<size_t>(h.record_bytes) * 4096u);
        uint64_t done = 0;
        while (done < h.points)
        {
            const uint64_t want = (h.points - done < 4096u) ? (h.points - done) : 4096u;
            if (std::fread(buf.data(), h.record_bytes, static_cast<size_t>(want), f) != want)
            {
                std::fclose(f);
                return "file shorter than POINTS says";
            }
            for (uint64_t i = 0; i < want; ++i)
            {
                const uint8_t *r = buf.data() + i * h.record_bytes;
                float x, y, z, in = 0.0f;
                std::memcpy(&x, r + h.off_x, 4);
                std::memcpy(&y, r + h.off_y, 4);
                std::memcpy(&z, r + h.off_z, 4);
                if (h.off_i >= 0)
                    std::memcpy(&in, r + h.off_i, 4);
                pcd_store(dst + (done + i) * stride_bytes, stride_bytes, x, y, z, in);
            }
            done += want;
        }
    }
    else
    {
        char line[8192];
        uint64_t done = 0;
        std::vector<float> vals(h.columns);
        while (done < h.points && std::fgets(line, sizeof(line), f))
        {
            char *p = line;
            uint32_t got = 0;
            while (got < h.columns)
            {
                char *end = nullptr;
                const float v = std::strtof(p, &end);
                if (end == p)
                    break;
                vals[got++] = v;
                p = end;
            }
            if (got == 0)
                continue; // blank line
            if (got < h.columns)
            {
                std::fclose(f);
                return "ascii record with too few columns";
            }
            pcd_store(dst + done * stride_bytes, stride_bytes, vals[h.col_x], vals[h.col_y], vals[h.col_z],
                      h.col_i >= 0 ? vals[h.col_i] : 0.0f);
            ++done;
        }
        if (done != h.points)
        {
            std::fclose(f);
            return "file shorter than POINTS says";
        }
    }
    std::fclose(f);
    return std::string();
}

} // namespace lb
