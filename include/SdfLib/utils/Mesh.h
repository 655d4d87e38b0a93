// sdflib::BoundingBox / sdflib::Mesh — drop-in mirror of include/SdfLib/utils/Mesh.h:16-106 of the reference for the
// B200-native library. Header-only; needs <glm/glm.hpp> exactly like the reference's public headers do.
// Mesh(std::string) (Mesh.h:76-79, assimp in the reference) reads Stanford PLY (ascii / binary_little_endian) and
// Wavefront OBJ directly: vertex positions and faces only, polygons triangulated as fans (aiProcess_Triangulate).
#ifndef SDFB200_SDFLIB_MESH_H
#define SDFB200_SDFLIB_MESH_H

#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
#include <glm/glm.hpp>

namespace sdflib
{
struct BoundingBox
{
    BoundingBox() : min(INFINITY), max(-INFINITY) {}
    BoundingBox(glm::vec3 min, glm::vec3 max) : min(min), max(max) {}
    glm::vec3 min;
    glm::vec3 max;

    glm::vec3 getSize() const { return max - min; }
    glm::vec3 getCenter() const { return min + 0.5f * getSize(); }
    void addMargin(float margin) { min -= glm::vec3(margin); max += glm::vec3(margin); }

    // Mesh.h:42-46
    float getDistance(glm::vec3 point) const
    {
        glm::vec3 q = glm::abs(point - getCenter()) - 0.5f * getSize();
        return glm::length(glm::max(q, glm::vec3(0.0f))) + glm::min(glm::max(q.x, glm::max(q.y, q.z)), 0.0f);
    }
};

class Mesh
{
public:
    Mesh() {}
    // src/utils/Mesh.cpp:34-42: the arrays are copied
    Mesh(glm::vec3* vertices, uint32_t numVertices, uint32_t* indices, uint32_t numIndices)
        : mVertices(vertices, vertices + numVertices), mIndices(indices, indices + numIndices)
    {
        computeBoundingBox();
    }

    // src/utils/Mesh.cpp:9-26. Throws std::runtime_error on unreadable / unsupported files (the reference logs and
    // leaves an empty mesh, which its builders then dereference).
    explicit Mesh(const std::string& filePath)
    {
        const size_t dot = filePath.find_last_of('.');
        std::string ext = dot == std::string::npos ? "" : filePath.substr(dot + 1);
        for (char& c : ext) c = char(std::tolower(static_cast<unsigned char>(c)));
        if (ext == "ply") loadPly(filePath);
        else if (ext == "obj") loadObj(filePath);
        else throw std::runtime_error("Mesh: unsupported model format '" + ext + "' (ply and obj are read)");
        if (mVertices.empty() || mIndices.empty()) throw std::runtime_error("Mesh: " + filePath + " holds no triangles");
        computeBoundingBox();
    }

    std::vector<glm::vec3>& getVertices() { return mVertices; }
    const std::vector<glm::vec3>& getVertices() const { return mVertices; }
    std::vector<uint32_t>& getIndices() { return mIndices; }
    const std::vector<uint32_t>& getIndices() const { return mIndices; }
    const BoundingBox& getBoundingBox() const { return mBBox; }

    void computeBoundingBox()   // src/utils/Mesh.cpp:66-76
    {
        glm::vec3 min(INFINITY), max(-INFINITY);
        for (const glm::vec3& v : mVertices)
        {
            min.x = glm::min(min.x, v.x); max.x = glm::max(max.x, v.x);
            min.y = glm::min(min.y, v.y); max.y = glm::max(max.y, v.y);
            min.z = glm::min(min.z, v.z); max.z = glm::max(max.z, v.z);
        }
        mBBox = BoundingBox(min, max);
    }

    void applyTransform(glm::mat4 trans)   // src/utils/Mesh.cpp:78-87
    {
        for (glm::vec3& v : mVertices) v = glm::vec3(trans * glm::vec4(v, 1.0f));
        computeBoundingBox();
    }

private:
    void addPolygon(const std::vector<uint32_t>& poly)
    {
        for (size_t k = 1; k + 1 < poly.size(); k++) { mIndices.push_back(poly[0]); mIndices.push_back(poly[k]); mIndices.push_back(poly[k + 1]); }
    }

    void loadObj(const std::string& path)
    {
        std::ifstream in(path);
        if (!in) throw std::runtime_error("Mesh: cannot open " + path);
        std::string line;
        std::vector<uint32_t> poly;
        while (std::getline(in, line))
        {
            std::istringstream ls(line);
            std::string tag;
            ls >> tag;
            if (tag == "v") { glm::vec3 v(0.0f); ls >> v.x >> v.y >> v.z; mVertices.push_back(v); }
            else if (tag == "f")
            {
                poly.clear();
                std::string ref;
                while (ls >> ref)
                {
                    const long i = std::strtol(ref.c_str(), nullptr, 10);   // "v", "v/vt", "v//vn", "v/vt/vn"
                    const long at = i > 0 ? i - 1 : long(mVertices.size()) + i;   // negative = relative to the vertices read so far
                    if (i == 0 || at < 0 || at > 0x7FFFFFFFL) throw std::runtime_error("Mesh: malformed face in " + path);
                    poly.push_back(uint32_t(at));
                }
                addPolygon(poly);
            }
        }
        for (uint32_t i : mIndices) if (i >= mVertices.size()) throw std::runtime_error("Mesh: face index out of range in " + path);
    }

    struct PlyProperty { std::string type, countType, name; bool list; };
    static size_t plySize(const std::string& t)
    {
        if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
        if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
        if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
        if (t == "double" || t == "float64") return 8;
        throw std::runtime_error("Mesh: unknown PLY type " + t);
    }
    static double plyRead(std::istream& in, const std::string& t, bool ascii)
    {
        if (ascii) { double v = 0; in >> v; return v; }
        char b[8] = {0};
        in.read(b, std::streamsize(plySize(t)));
        if (t == "char" || t == "int8") { int8_t v; std::memcpy(&v, b, 1); return v; }
        if (t == "uchar" || t == "uint8") { uint8_t v; std::memcpy(&v, b, 1); return v; }
        if (t == "short" || t == "int16") { int16_t v; std::memcpy(&v, b, 2); return v; }
        if (t == "ushort" || t == "uint16") { uint16_t v; std::memcpy(&v, b, 2); return v; }
        if (t == "int" || t == "int32") { int32_t v; std::memcpy(&v, b, 4); return v; }
        if (t == "uint" || t == "uint32") { uint32_t v; std::memcpy(&v, b, 4); return v; }
        if (t == "float" || t == "float32") { float v; std::memcpy(&v, b, 4); return v; }
        double v; std::memcpy(&v, b, 8); return v;
    }

    void loadPly(const std::string& path)
    {
        std::ifstream in(path, std::ios::binary);
        if (!in) throw std::runtime_error("Mesh: cannot open " + path);
        std::string line;
        std::getline(in, line);
        if (line.substr(0, 3) != "ply") throw std::runtime_error("Mesh: " + path + " is not a PLY file");
        bool ascii = true;
        struct Element { std::string name; size_t count; std::vector<PlyProperty> props; };
        std::vector<Element> elements;
        while (std::getline(in, line))
        {
            if (!line.empty() && line.back() == '\r') line.pop_back();
            std::istringstream ls(line);
            std::string tag;
            ls >> tag;
            if (tag == "format")
            {
                std::string f; ls >> f;
                if (f == "ascii") ascii = true;
                else if (f == "binary_little_endian") ascii = false;
                else throw std::runtime_error("Mesh: PLY format " + f + " is not supported");
            }
            else if (tag == "element") { Element e; ls >> e.name >> e.count; elements.push_back(e); }
            else if (tag == "property")
            {
                if (elements.empty()) throw std::runtime_error("Mesh: PLY property before any element");
                PlyProperty p; p.list = false;
                ls >> p.type;
                if (p.type == "list") { p.list = true; ls >> p.countType >> p.type; }
                ls >> p.name;
                elements.back().props.push_back(p);
            }
            else if (tag == "end_header") break;
        }
        std::vector<uint32_t> poly;
        for (const Element& e : elements)
        {
            if (e.props.empty()) continue;   // nothing to read per entry: a huge count must not spin
            if (e.count > 0xFFFFFFFFull) throw std::runtime_error("Mesh: " + path + " declares an implausible element count");
            for (size_t i = 0; i < e.count; i++)
            {
                glm::vec3 v(0.0f);
                for (const PlyProperty& p : e.props)
                {
                    if (p.list)
                    {
                        // a damaged file must end in an exception, not in a loop over 2^32 entries or a wild cast
                        const double count = plyRead(in, p.countType, ascii);
                        if (!in || !(count >= 0.0 && count <= 65536.0)) throw std::runtime_error("Mesh: " + path + " holds a malformed list");
                        const size_t n = size_t(count);
                        poly.clear();
                        for (size_t k = 0; k < n; k++)
                        {
                            const double x = plyRead(in, p.type, ascii);
                            if (!in || !(x >= 0.0 && x <= 4294967295.0)) throw std::runtime_error("Mesh: " + path + " holds a malformed list");
                            poly.push_back(uint32_t(x));
                        }
                        if (e.name == "face" && (p.name == "vertex_indices" || p.name == "vertex_index")) addPolygon(poly);
                    }
                    else
                    {
                        const double x = plyRead(in, p.type, ascii);
                        if (p.name == "x") v.x = float(x); else if (p.name == "y") v.y = float(x); else if (p.name == "z") v.z = float(x);
                    }
                }
                if (e.name == "vertex") mVertices.push_back(v);
                if (!in) throw std::runtime_error("Mesh: " + path + " ends early");
            }
        }
        for (uint32_t i : mIndices) if (i >= mVertices.size()) throw std::runtime_error("Mesh: face index out of range in " + path);
    }

    std::vector<glm::vec3> mVertices;
    std::vector<uint32_t> mIndices;
    BoundingBox mBBox;
};
}

#endif
