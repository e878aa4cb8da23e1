// fastnoise2_b200 -- symbol visibility (same macro name as the reference's Utility/Export.h:4-15)
#ifndef FASTNOISE_B200_EXPORT_H
#define FASTNOISE_B200_EXPORT_H

#if defined( _WIN32 ) && !defined( FASTNOISE_STATIC_LIB )
#  ifdef FASTNOISE_EXPORT
#    define FASTNOISE_API __declspec( dllexport )
#  else
#    define FASTNOISE_API __declspec( dllimport )
#  endif
#else
#  define FASTNOISE_API __attribute__(( visibility( "default" ) ))
#endif

#endif
