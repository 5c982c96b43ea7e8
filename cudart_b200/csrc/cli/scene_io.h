// cudart_b200/csrc/cli/scene_io.h - host-side scene loading for the drop-in CLI, in the repo's own C++.
//
// Reproduces the OBSERVABLE behaviour of the reference's loading surface so that `./CudaRT -f <scene> ...`
// consumes the same files and hands the renderer the same arrays:
//   EnvFile            <- EnvLoader            (context/env_loader.cpp:14-49)
//   load_camera_json   <- CameraJsonLoader     (context/json_loader.cpp:14-64, picojson)
//   load_obj           <- objl::Loader          (include/obj_load.h:441-723, 1016-1186; triangulated faces only)
//   load_png_as_float3 <- TextureStore::load    (context/texture_store.cpp:18-35, stb_image stbi_loadf)
//   assemble_scene     <- Scene::loadTextures / loadTriangles / constructBvh / constructLightList
//                         (geometry/scene.cpp:55-203), BVH + lights through the C ABI (crt_build_bvh/_lights)
//   write_png_rgba8    <- Launcher::saveToImage (context/launcher.cpp:13-21, stb_image_write)
// including the quirks that shape the data.
#pragma once
#include <cstdint>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../host_scene.h"

namespace crtcli {

struct EnvFile {
    explicit EnvFile(const std::string& path);          // throws std::runtime_error when the file cannot be opened
    std::string meshes, camera, textures;
};

crt::RefCamera load_camera_json(const std::string& path, unsigned width, unsigned height);

struct ObjMaterial {                                    // objl::Material (obj_load.h:149-194)
    std::string name = "not-initialized";
    float Ka[3] = {0, 0, 0}, Kd[3] = {0, 0, 0}, Ks[3] = {0, 0, 0};
    float Ns = 0.f, Ni = 0.f, diffuse = 0.f;
    std::string map_Kd, type = "LAMBERT";
};
struct ObjVertex { float p[3] = {0, 0, 0}, n[3] = {0, 0, 0}, t[2] = {0, 0}; };
struct ObjMesh {
    std::string name;
    std::vector<ObjVertex> vertices;                    // three per face (vertex soup, D25)
    ObjMaterial material;
};
struct ObjFile {
    std::vector<ObjMesh> meshes;
    std::vector<ObjMaterial> materials;                 // LoadedMaterials, file order
    std::vector<ObjVertex> all_vertices;                // LoadedVertices (what the BVH builder boxes)
};
ObjFile load_obj(const std::string& path);

struct Texture { unsigned w = 0, h = 0; std::vector<float> rgb; };
Texture load_png_as_float3(const std::string& path);
void write_png_rgba8(const std::string& path, const uint8_t* rgba, unsigned w, unsigned h);

// Everything crt_scene_desc points at, owned.
struct SceneArrays {
    std::vector<crt::RefTriangle> tris;
    std::vector<crt::RefNode> nodes;
    std::vector<float> mat_floats;                      // n_mats x 12
    std::vector<int32_t> mat_indices;                   // n_mats x 2
    std::vector<uint32_t> lights;
    float lights_area = 0.f;
    crt::RefCamera camera{};
    std::vector<float> tex_data;                        // float3, flattened + padded
    std::vector<uint32_t> tex_dims, tex_offsets;
    uint32_t n_mats = 0, n_tex = 0;
    std::vector<std::pair<std::string, size_t>> mesh_report;   // (name, vertices) for the CLI's per-mesh printout
};
SceneArrays assemble_scene(const ObjFile& obj, const crt::RefCamera& cam, const std::string& textures_dir);

// CRTS container (cudart_b200/scenefile.py)
void save_crts(const std::string& path, const SceneArrays& s, unsigned width, unsigned height);

}  // namespace crtcli
