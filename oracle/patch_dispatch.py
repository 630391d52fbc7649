#!/usr/bin/env python
"""Writes patched COPIES of the two reference files a maintainer touches to register the B200
back-end into the git-ignored build directory oracle/_ref/patched/:
  src/RayTraceImage.cpp  one extern block, two `else if` branches next to the "cuda" ones
                         (:47-75, :389-405) and the two direct entries ("b200-direct",
                         "b200-multigpu") ahead of the host ray list (:277);
  src/CreateImage.cpp    the method name in the driver's default list and GPU warm-up (:90-132).
Nothing else of the reference is touched; the copies never enter the repository.  See
INTEGRATION.md for the same change as a diff."""
import os
import sys

src, dst = sys.argv[1], sys.argv[2]
s = open(src).read()
if os.path.basename(src) == "CreateImage.cpp":
    # driver registration: default method list + GPU warm-up are keyed on literal method names
    # (src/CreateImage.cpp:90-132)
    a1 = '        methods.push_back( "threads" );\n'
    assert s.count(a1) == 1
    s = s.replace(a1, a1 + '        methods.push_back( "b200" );\n')
    a2 = '        auto index = std::find(methods.begin(),methods.end(),"Cuda-MultiGPU");\n'
    assert s.count(a2) == 1
    s = s.replace(a2, '        auto index = std::find(methods.begin(),methods.end(),"b200");\n'
                      '        if ( index == methods.end() )\n'
                      '            index = std::find(methods.begin(),methods.end(),"Cuda-MultiGPU");\n')
    os.makedirs(os.path.dirname(dst), exist_ok=True)
    open(dst, "w").write(s)
    sys.exit(0)
extern = '''
// ---- B200 back-end (raytrace-miniapp_b200/host/RayTraceImageB200.cpp) ----
extern void RayTraceImageB200Loop( int N, const RayTrace::EUV_beam_struct& euv_beam, const RayTrace::ray_gain_struct *gain,
    const RayTrace::ray_seed_struct *seed, int method, const std::vector<ray_struct> &rays,
    double scale, double *image, double *I_ang, unsigned int &failure_code,
    std::vector<ray_struct> &failed_rays );
extern void RayTraceImageB200SetDevice( int device );
extern void RayTraceImageB200Direct( const RayTrace::create_image_struct *info, double *image, double *I_ang );
extern void RayTraceImageB200MultiDirect( const RayTrace::create_image_struct *info, double *image, double *I_ang );
extern "C" int rtb200_device_count( void );
'''
anchor = "/**********************************************************************\n* Call RayTraceImage function from a thread loop"
assert s.count(anchor) == 1
s = s.replace(anchor, extern + "\n" + anchor)
branch = '''    } else if ( compute_method == "b200" ) {
        RayTraceImageB200Loop( N, std::ref(*info->euv_beam), info->gain, info->seed,
            method, rays, scale, image, I_ang, failure_code, failed_rays );
    } else if ( compute_method == "b200-threadloop" ) {
        RayTraceImageThreadLoop( rtb200_device_count(), RayTraceImageB200Loop, RayTraceImageB200SetDevice,
            N, std::ref(*info->euv_beam), info->gain, info->seed,
            method, rays, scale, image, I_ang, failure_code, failed_rays );
'''
# full-speed entry, before the host ray list is built (:277)
direct = '''    {
        std::string m2 = compute_method;
        std::transform( m2.begin(), m2.end(), m2.begin(), ::tolower );
        if ( m2 == "b200-direct" ) {
            RayTraceImageB200Direct( info, image, I_ang );
            PROFILE_STOP( "create_image" );
            return;
        }
        if ( m2 == "b200-multigpu" ) {
            RayTraceImageB200MultiDirect( info, image, I_ang );
            PROFILE_STOP( "create_image" );
            return;
        }
    }

'''
anchor3 = "    // Create a list of rays to propagate\n"
assert s.count(anchor3) == 1
s = s.replace(anchor3, direct + anchor3)
anchor2 = '    } else if ( compute_method == "cpu" ) {'
assert s.count(anchor2) == 1
s = s.replace(anchor2, branch + anchor2)
os.makedirs(os.path.dirname(dst), exist_ok=True)
open(dst, "w").write(s)
