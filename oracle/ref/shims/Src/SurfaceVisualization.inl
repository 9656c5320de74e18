// Data-only stand-in for the reference's GLUT viewer (include/Src/SurfaceVisualization.inl,
// include/Misha/Visualization.h): exactly the members OpticalFlow.cpp touches, no GL.
// Test infrastructure only.
#ifndef MOF_SHIM_SURFACE_VISUALIZATION
#define MOF_SHIM_SURFACE_VISUALIZATION
#include <string.h>
#include <vector>
#include <Misha/Geometry.h>
#include <Misha/Image.h>
struct Visualization
{
	int screenWidth, screenHeight;
	struct KeyboardCallBack
	{
		char key;
		void (*callBackFunction)(Visualization*, const char*);
		KeyboardCallBack(Visualization*, char key, const char*, void (*f)(Visualization*, const char*)) : key(key), callBackFunction(f) {}
		KeyboardCallBack(Visualization*, char key, const char*, const char*, void (*f)(Visualization*, const char*)) : key(key), callBackFunction(f) {}
	};
	std::vector<KeyboardCallBack> callBacks;
	std::vector<char*> info;
	Visualization(void) : screenWidth(512), screenHeight(512) {}
	void Idle(void) {}
	void KeyboardFunc(unsigned char, int, int) {}
	void SpecialFunc(int, int, int) {}
	void Display(void) {}
	void Reshape(int, int) {}
	void MouseFunc(int, int, int, int) {}
	void MotionFunc(int, int) {}
};
struct SurfaceVisualization : public Visualization
{
	bool useTexture;
	unsigned char* texture;
	int textureWidth, textureHeight;
	std::vector<Point2D<float> > textureCoordinates;
	std::vector<TriangleIndex> triangles;
	std::vector<Point3D<float> > vertices, colors, vectorField;
	SurfaceVisualization(void) : useTexture(false), texture(NULL), textureWidth(0), textureHeight(0) {}
	void updateTextureBuffer(bool = false) {}
	void updateMesh(bool) {}
};
#endif
