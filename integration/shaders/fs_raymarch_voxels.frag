#version 460
// fs_raymarch_voxels.frag — voxel-DDA variant of the reference's shaders/fs_raymarch.frag (SURVEY.md §8(f).1).
//
// Same stage interface as the reference shader (locations 0-2 from shaders/fs_raymarch.vert:30-37, which is reused
// unchanged), same camera (fs_raymarch.frag:67-75) and the same lighting (fs_raymarch.frag:49-55: light (2,5,3),
// direction = normalize(p - light) as shipped, max(0.05, n.l)); map_the_world's analytic sphere is replaced by an
// Amanatides-Woo walk through the uint8 material grid that libfs3d exports (fs3d_volume_export_fd), and the colour
// comes from a 256-entry palette — the reference's own, never-used colors[256] (renderer.cpp:136-393).
// This file is the GLSL twin of fallingsand3d_b200/csrc/raymarch.cuh (mode FS3D_RM_VOXELS): statement for statement
// the same walk, so the offscreen CUDA image (pixel-exact against oracle/fs3d_raymarch_oracle.c and
// oracle/oracle_np_dda.py) is the golden for it.  It cannot be compiled in this image (no glslang); it is text for the
// maintainer, to be compiled by the reference's own rule (CMakeLists.txt:22-37 globs shaders/*.frag).
//
// Set 2 (the material set, materials.cpp:182-196 reserve sets 0 and 1):
//   binding 0  uniform VolumeInfo   grid size, z-range held by this buffer, voxel edge, half extents
//   binding 1  readonly buffer      the cells, one byte each, x fastest: 4 cells per uint
//   binding 2  uniform Palette      256 x vec4

layout (location = 0) in vec3 inFragOrigin;
layout (location = 1) in vec2 inUV;
layout (location = 2) in float inAspect;

layout (location = 0) out vec4 outFragColor;

layout (set = 2, binding = 0) uniform VolumeInfo {
	uvec4 grid;        // nx, ny, nz, first cell's byte offset / 4 inside the buffer (fs3d_export.first_cell_offset / 4)
	uvec4 zrange;      // z0, z1 of the planes this buffer holds, 0, 0
	vec4  extent;      // ex, ey, ez (half extents of the box), h (voxel edge = 1 / max(nx, ny, nz))
} volume;

layout (std430, set = 2, binding = 1) readonly buffer Cells {
	uint cells[];
} cellBuffer;

layout (set = 2, binding = 2) uniform Palette {
	vec4 colors[256];
} palette;

uint voxel_at(ivec3 i) {
	// grid +y is world -y (this camera's screen-down is world +y): the caller passes the flipped row
	if (uint(i.z) < volume.zrange.x || uint(i.z) >= volume.zrange.y) return 0u;
	uint idx = uint(i.x) + volume.grid.x * (uint(i.y) + volume.grid.y * (uint(i.z) - volume.zrange.x));
	uint word = cellBuffer.cells[volume.grid.w + (idx >> 2)];
	return (word >> (8u * (idx & 3u))) & 0xFFu;
}

float diffuse_at(vec3 p, vec3 n) {
	vec3 light_pos = vec3(2.0, 5.0, 3.0);
	vec3 direction_to_light = normalize(p - light_pos);      // as shipped (fs_raymarch.frag:52)
	return max(0.05, dot(n, direction_to_light));
}

void main() {
	vec2 uv = inUV * 2.0 - 1.0;
	uv.y /= inAspect;
	vec3 o = inFragOrigin;
	vec3 d = normalize(vec3(uv, 1.0));
	vec3 e = volume.extent.xyz;
	float h = volume.extent.w;
	ivec3 n = ivec3(volume.grid.xyz);

	vec3 color = vec3(0.0);

	// slab test against the box [-e, e]
	float tmin = 0.0, tmax = 1.0 / 0.0;
	bool miss = false;
	for (int a = 0; a < 3; ++a) {
		if (d[a] != 0.0) {
			float t0 = (-e[a] - o[a]) / d[a], t1 = (e[a] - o[a]) / d[a];
			if (t0 > t1) { float s = t0; t0 = t1; t1 = s; }
			tmin = max(tmin, t0);
			tmax = min(tmax, t1);
		} else if (o[a] < -e[a] || o[a] > e[a]) {
			miss = true;
		}
	}

	if (!miss && tmin <= tmax) {
		// crossing times come from the integer boundary index, ((k h - e) - o) * (1 / d): never accumulated (raymarch.cuh)
		ivec3 idx, stp;
		vec3 tnext, rcp;
		int last_axis = -1;
		float best = -1.0;
		for (int a = 0; a < 3; ++a) {
			float pos = o[a] + tmin * d[a];
			int i = clamp(int(floor((pos + e[a]) / h)), 0, n[a] - 1);
			idx[a] = i;
			stp[a] = d[a] > 0.0 ? 1 : (d[a] < 0.0 ? -1 : 0);
			rcp[a] = stp[a] != 0 ? 1.0 / d[a] : 0.0;
			tnext[a] = stp[a] != 0 ? ((float(i + (stp[a] > 0 ? 1 : 0)) * h - e[a]) - o[a]) * rcp[a] : 1.0 / 0.0;
			if (d[a] != 0.0) {              // the face the ray entered through: the axis whose slab entry time is tmin
				float t0 = (-e[a] - o[a]) / d[a], t1 = (e[a] - o[a]) / d[a];
				float tn = min(t0, t1);
				if (tn == tmin && tn > best) { best = tn; last_axis = a; }
			}
		}
		float t = tmin;
		int max_steps = n.x + n.y + n.z + 3;
		for (int s = 0; s < max_steps; ++s) {
			uint m = voxel_at(ivec3(idx.x, n.y - 1 - idx.y, idx.z));
			if (m != 0u) {
				vec3 nrm = vec3(0.0);
				if (last_axis >= 0) nrm[last_axis] = stp[last_axis] > 0 ? -1.0 : 1.0;
				color = palette.colors[m].rgb * diffuse_at(o + t * d, nrm);
				break;
			}
			int a = 0;
			if (tnext[1] < tnext[a]) a = 1;
			if (tnext[2] < tnext[a]) a = 2;
			t = tnext[a];
			idx[a] += stp[a];
			if (idx[a] < 0 || idx[a] >= n[a]) break;
			tnext[a] = ((float(idx[a] + (stp[a] > 0 ? 1 : 0)) * h - e[a]) - o[a]) * rcp[a];
			last_axis = a;
		}
	}

	outFragColor = vec4(color, 1.0);
}
