// TEST INFRASTRUCTURE ONLY.  The reference's two ASCII-STL readers are flex sources
// (src/triSurface/triSurface/interfaces/STL/readSTLASCII.L, src/surfMesh/surfaceFormats/stl/STLsurfaceFormatASCII.L)
// and this image has no flex.  The icoFoam oracle (oracle/build_app.py) never reads an STL file, so the two entry
// points are provided as stubs that fail loudly.
#include "triSurface.H"
#include "STLsurfaceFormatCore.H"
#include "error.H"

bool Foam::triSurface::readSTLASCII(const fileName& f)
{
    FatalErrorInFunction
        << "ASCII STL reading is not available in the oracle build (no flex): " << f
        << exit(FatalError);
    return false;
}

bool Foam::fileFormats::STLsurfaceFormatCore::readASCII(istream&, const off_t)
{
    FatalErrorInFunction
        << "ASCII STL reading is not available in the oracle build (no flex)"
        << exit(FatalError);
    return false;
}
