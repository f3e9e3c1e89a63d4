# Applied to a COPY of bvh/bvh.cpp at build time (oracle/Makefile; the copy lives under
# oracle/_ref/patched/, git-ignored - no reference source enters the repository).
# It is the one-line hook of INTEGRATION.md section 2: BVH::create asks a registered
# device plug-in for layouts this tree does not know before its own switch.
/^BVH \*BVH::create(const BVHParams &params,$/i\
BVH *bvh_layout_hook_create(const BVHParams &params,\
                            const vector<Geometry *> &geometry,\
                            const vector<Object *> &objects);\

/^  switch (params.bvh_layout) {$/i\
  if (BVH *plugin_bvh = bvh_layout_hook_create(params, geometry, objects))\
    return plugin_bvh;
