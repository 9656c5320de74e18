// TEST INFRASTRUCTURE ONLY — never linked into or called by the product.
//
// Entry point of oracle/_ref/OpticalFlow_ref: the UNMODIFIED reference solver
// (OpticalFlow/OpticalFlow.cpp and the include/ headers, compiled where they lie under
// /root/reference by oracle/ref/build_ref.sh) behind its own command line, plus an optional
// `--tap <dir>` that dumps intermediate state as .npy files for the parity tests.
//
// Without --tap this runs exactly the reference's headless path
// (main -> _main<double,3> -> Init -> IterativeOptimization, OpticalFlow.cpp:1096-1116,
// 1059-1094, 1036-1056) and is what bench.py times as the CPU baseline.
// With --tap it calls the same reference functions in the same order as UpdateFlow
// (OpticalFlow.cpp:424-474) and VectorField::UpdateOpticalFlow (VectorField.h:46-104),
// writing what passes between them. No reference arithmetic is re-implemented here except the
// tap-only recomputation of (A, b, x) through the reference's own operators.
#include <omp.h>
#include <Misha/CmdLineParser.h>
extern cmdLineReadable Verbose;  // VectorField.h:54 uses it before OpticalFlow.cpp:62 defines it

#define main reference_main
#include "OpticalFlow.gen.cpp"
#undef main

#include <stdint.h>
#include <string>
#include <sys/stat.h>

namespace tap {
std::string dir;

void npy(const std::string& name, const char* descr, const void* data, size_t rows, size_t cols, size_t elem)
{
	std::string path = dir + "/" + name + ".npy";
	FILE* fp = fopen(path.c_str(), "wb");
	if (!fp) fprintf(stderr, "[ERROR] cannot write %s\n", path.c_str()), exit(1);
	char shape[64];
	if (cols) sprintf(shape, "(%zu, %zu)", rows, cols);
	else sprintf(shape, "(%zu,)", rows);
	char dict[256];
	int n = sprintf(dict, "{'descr': '%s', 'fortran_order': False, 'shape': %s, }", descr, shape);
	int total = 10 + n + 1;
	int pad = (64 - total % 64) % 64;
	unsigned char head[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, 0, 0};
	uint16_t hlen = (uint16_t)(n + pad + 1);
	head[8] = (unsigned char)(hlen & 0xff), head[9] = (unsigned char)(hlen >> 8);
	fwrite(head, 1, 10, fp);
	fwrite(dict, 1, n, fp);
	for (int i = 0; i < pad; i++) fputc(' ', fp);
	fputc('\n', fp);
	fwrite(data, elem, rows * (cols ? cols : 1), fp);
	fclose(fp);
}
void f64(const std::string& name, const double* d, size_t rows, size_t cols = 0) { npy(name, "<f8", d, rows, cols, 8); }
void i32(const std::string& name, const int* d, size_t rows, size_t cols = 0) { npy(name, "<i4", d, rows, cols, 4); }

// CSR with each row's columns sorted ascending (the reference's SpGEMM leaves them in
// unordered_map order, SparseMatrix.inl:372-388); keepOrder dumps rows as stored.
void csr(const std::string& name, const SparseMatrix< double , int >& M, bool keepOrder = false)
{
	std::vector< int > rowptr(M.rows + 1, 0), col;
	std::vector< double > val;
	for (size_t i = 0; i < M.rows; i++)
	{
		std::vector< std::pair< int , double > > row;
		for (size_t j = 0; j < M.rowSizes[i]; j++) row.push_back(std::make_pair(M[i][j].N, M[i][j].Value));
		if (!keepOrder) std::sort(row.begin(), row.end(), [](const std::pair< int , double >& a, const std::pair< int , double >& b){ return a.first < b.first; });
		for (size_t j = 0; j < row.size(); j++) col.push_back(row[j].first), val.push_back(row[j].second);
		rowptr[i + 1] = (int)col.size();
	}
	i32(name + ".rowptr", rowptr.data(), rowptr.size());
	i32(name + ".col", col.data(), col.size());
	f64(name + ".val", val.data(), val.size());
}
template< int Channels >
void signal(const std::string& name, const std::vector< Point< double , Channels > >& s)
{
	std::vector< double > flat(s.size() * Channels);
	for (size_t i = 0; i < s.size(); i++) for (int c = 0; c < Channels; c++) flat[i * Channels + c] = s[i][c];
	f64(name, flat.data(), s.size(), Channels);
}
void points3(const std::string& name, const std::vector< Point3D< double > >& s)
{
	std::vector< double > flat(s.size() * 3);
	for (size_t i = 0; i < s.size(); i++) for (int c = 0; c < 3; c++) flat[i * 3 + c] = s[i][c];
	f64(name, flat.data(), s.size(), 3);
}
}  // namespace tap

typedef WhitneyFlowViewer< double , 3 > Viewer;

static void TapInit(void)
{
	FlowData< double , 3 >& fd = Viewer::flowData;
	int vCount = (int)fd.vertices.size(), tCount = (int)fd.triangles.size();
	tap::points3("vertices", fd.vertices);
	{
		std::vector< int > tri(3 * tCount);
		for (int t = 0; t < tCount; t++) for (int j = 0; j < 3; j++) tri[3 * t + j] = fd.triangles[t][j];
		tap::i32("triangles", tri.data(), tCount, 3);
	}
	{
		std::vector< double > g(4 * tCount);
		for (int t = 0; t < tCount; t++) for (int c = 0; c < 2; c++) for (int r = 0; r < 2; r++) g[4 * t + 2 * c + r] = fd.mesh.g[t](c, r);
		tap::f64("g", g.data(), tCount, 4);
		tap::f64("triangleArea", fd.triangleArea.data(), tCount);
	}
	{
		std::vector< int > opp(3 * tCount);
		std::vector< double > lin(12 * tCount), cst(6 * tCount);
		for (int e = 0; e < 3 * tCount; e++)
		{
			opp[e] = fd.edges[e].oppositeEdge;
			for (int c = 0; c < 2; c++) for (int r = 0; r < 2; r++) lin[4 * e + 2 * c + r] = fd.edges[e].xForm.linear(c, r);
			for (int c = 0; c < 2; c++) cst[2 * e + c] = fd.edges[e].xForm.constant[c];
		}
		tap::i32("oppositeEdge", opp.data(), opp.size());
		tap::f64("xform_linear", lin.data(), 3 * tCount, 4);  // column-major: (c,r) at 2c+r
		tap::f64("xform_constant", cst.data(), 3 * tCount, 2);
	}
	tap::csr("sMass", fd.sMass, true);
	tap::csr("sStiffness", fd.sStiffness, true);
	tap::signal< 3 >("signals0", fd.signals[0]);
	tap::signal< 3 >("signals1", fd.signals[1]);
	if (VectorFieldMode.value == WHITNEY_VECTOR_FIELD)
	{
		WhitneyVectorField< double >* w = (WhitneyVectorField< double >*)Viewer::vf;
		std::vector< int > pos(w->positiveOrientedEdge.size());
		for (size_t i = 0; i < pos.size(); i++) pos[i] = w->positiveOrientedEdge[i] ? 1 : 0;
		tap::i32("reducedEdgeIndex", w->reducedEdgeIndex.data(), w->reducedEdgeIndex.size());
		tap::i32("expandedEdgeIndex", w->expandedEdgeIndex.data(), w->expandedEdgeIndex.size());
		tap::i32("positiveOrientedEdge", pos.data(), pos.size());
	}
	tap::csr("prolongation", Viewer::vf->prolongationOperator, true);
	tap::csr("smoothOperator", Viewer::vf->smoothOperator);
	if (Viewer::processTexture)
	{
		InputTextureData< double >& td = Viewer::inputTextureData;
		int n = td.tWidth * td.tHeight;
		std::vector< int > tIdx(n);
		std::vector< double > p(2 * n, 0.), tt(td.triangleTextures.size() * 2);
		for (int i = 0; i < n; i++)
		{
			tIdx[i] = td.textureSource[i].tIdx;
			if (tIdx[i] != -1) p[2 * i] = td.textureSource[i].p[0], p[2 * i + 1] = td.textureSource[i].p[1];
		}
		for (size_t i = 0; i < td.triangleTextures.size(); i++) tt[2 * i] = td.triangleTextures[i][0], tt[2 * i + 1] = td.triangleTextures[i][1];
		int wh[2] = {td.tWidth, td.tHeight};
		tap::i32("texture_size", wh, 2);
		tap::i32("textureSource_tIdx", tIdx.data(), n);
		tap::f64("textureSource_p", p.data(), n, 2);
		tap::f64("triangleTextures", tt.data(), td.triangleTextures.size(), 2);
		for (int s = 0; s < 2; s++)
		{
			std::vector< int > tex(3 * n);
			for (int i = 0; i < 3 * n; i++) tex[i] = td.textures[s][i];
			char name[64];
			sprintf(name, "texture%d", s);
			tap::i32(name, tex.data(), n, 3);
		}
	}
	else
	{
		tap::points3("colors0", Viewer::inputGeometryData.colors[0]);
		tap::points3("colors1", Viewer::inputGeometryData.colors[1]);
	}
}

// Same calls, same order as UpdateFlow (OpticalFlow.cpp:424-474), tapping what flows between them.
static void TappedUpdateFlow(int iter, double sWeight, double vfWeight)
{
	FlowData< double , 3 >& fd = Viewer::flowData;
	VectorField< double >* vf = Viewer::vf;
	char pre[64];
	sprintf(pre, "it%02d.", iter);
	std::string P(pre);

	std::vector< Point< double , 3 > > smoothed[2], resampled[2];
	if (sWeight) for (int s = 0; s < 2; s++) fd.smoothSignal(fd.signals[s], smoothed[s], sWeight);
	tap::signal< 3 >(P + "smoothed0", smoothed[0]), tap::signal< 3 >(P + "smoothed1", smoothed[1]);
	for (int s = 0; s < 2; s++) ResampleSignal(fd.mesh, (ConstPointer(Point2D< double >))GetPointer(fd.tFlowField), (ConstPointer(FEM::EdgeXForm< double >))fd.edges, smoothed[s], resampled[s], (double)(s == 0 ? -0.5 : 0.5), Threads.value);
	tap::signal< 3 >(P + "resampled0", resampled[0]), tap::signal< 3 >(P + "resampled1", resampled[1]);

	SparseMatrix< double , int > dataTerm;
	std::vector< double > rhs;
	SetDataTerm(fd.triangles, fd.triangleArea, resampled, dataTerm, rhs);
	{
		int tCount = (int)fd.triangles.size();
		std::vector< double > D(4 * tCount);
		for (int t = 0; t < tCount; t++) for (int k = 0; k < 2; k++) for (int l = 0; l < 2; l++) D[4 * t + 2 * k + l] = dataTerm[2 * t + k][l].Value;
		tap::f64(P + "dataTerm", D.data(), tCount, 4);
		tap::f64(P + "rhs", rhs.data(), rhs.size());
	}
	// Tap-only recomputation of the system VectorField::UpdateOpticalFlow builds (VectorField.h:51-67,85),
	// through the reference's own operators.
	{
		SparseMatrix< double , int > D = vf->restrictionOperator * dataTerm * vf->prolongationOperator;
		std::vector< double > b(vf->coeffs.size()), x(vf->coeffs.size(), 0.);
		vf->restrictionOperator.Multiply(GetPointer(rhs), GetPointer(b));
		double scale = 1. / sqrt(D.SquareNorm());
		D *= scale;
		for (size_t i = 0; i < b.size(); i++) b[i] *= scale;
		SparseMatrix< double , int > A = D + vf->smoothOperator * vfWeight;
		EigenCholeskySolverLDLt solver(A);
		solver.solve(GetPointer(b), GetPointer(x));
		tap::csr(P + "A", A);
		tap::f64(P + "b", b.data(), b.size());
		tap::f64(P + "x", x.data(), x.size());
		tap::f64(P + "scale", &scale, 1);
	}
	vf->UpdateOpticalFlow(dataTerm, rhs, vfWeight, fd.tFlowField);
	tap::f64(P + "coeffs", vf->coeffs.data(), vf->coeffs.size());
	{
		std::vector< double > tf(2 * fd.tFlowField.size());
		for (size_t t = 0; t < fd.tFlowField.size(); t++) tf[2 * t] = fd.tFlowField[t][0], tf[2 * t + 1] = fd.tFlowField[t][1];
		tap::f64(P + "tFlowField", tf.data(), fd.tFlowField.size(), 2);
	}
}

int main(int argc, char* argv[])
{
	// Strip our own flags before the reference's parser sees the rest.
	std::vector< char* > args;
	bool timeOnly = false;
	for (int i = 0; i < argc; i++)
	{
		if (!strcmp(argv[i], "--tap") && i + 1 < argc) tap::dir = argv[++i];
		else if (!strcmp(argv[i], "--stageTimes")) timeOnly = true;
		else args.push_back(argv[i]);
	}
	(void)timeOnly;
	if (tap::dir.empty()) return reference_main((int)args.size(), args.data());

	mkdir(tap::dir.c_str(), 0755);
	// From here: main (OpticalFlow.cpp:1096-1116) and _main<double,3> (:1059-1074), with taps.
	cmdLineParse((int)args.size() - 1, args.data() + 1, params);
	if (!In.set || !Out.set) { ShowUsage(args[0]); return EXIT_FAILURE; }
	DoGWeight.value = std::min< float >(1.f, std::max< float >(0.f, DoGWeight.value));
	if (DoGWeight.value > 0 && DoGWeight.value < 1) { fprintf(stderr, "[ERROR] --tap supports the 3-channel path only\n"); return EXIT_FAILURE; }
	Viewer::scalarSmoothWeight = (double)ScalarSmoothWeight.value;
	if (VectorFieldSmoothWeight.set) Viewer::vectorFieldSmoothWeight = (double)VectorFieldSmoothWeight.value;
	else
	{
		if (VectorFieldMode.value == WHITNEY_VECTOR_FIELD) Viewer::vectorFieldSmoothWeight = 3e-6;
		if (VectorFieldMode.value == CONFORMAL_VECTOR_FIELD) Viewer::vectorFieldSmoothWeight = 5e-7;
		if (VectorFieldMode.value == CONNECTION_VECTOR_FIELD) Viewer::vectorFieldSmoothWeight = 1e4;
	}
	if (!Viewer::Init()) return 0;
	TapInit();
	// IterativeOptimization (OpticalFlow.cpp:1036-1056) with the tapped UpdateFlow.
	for (int i = 0; i < Levels.value; i++)
	{
		TappedUpdateFlow(i, Viewer::scalarSmoothWeight, Viewer::vectorFieldSmoothWeight);
		Viewer::scalarSmoothWeight *= ScalarWeightMultiplier.value;
		Viewer::vectorFieldSmoothWeight = Viewer::vectorFieldSmoothWeight * VectorFieldWeightMultiplier.value > VectorFieldSmoothWeightThreshold.value ? Viewer::vectorFieldSmoothWeight * VectorFieldWeightMultiplier.value : Viewer::vectorFieldSmoothWeight;
	}
	FlowData< double , 3 >& fd = Viewer::flowData;
	if (Viewer::processTexture)
	{
		InputTextureData< double >& td = Viewer::inputTextureData;
		td.flow(fd, 0.5, Viewer::inputAdvectedTexture, Threads.value);
		int n = td.tWidth * td.tHeight;
		for (int s = 0; s < 2; s++)
		{
			std::vector< double > flat(3 * n);
			for (int i = 0; i < n; i++) for (int c = 0; c < 3; c++) flat[3 * i + c] = Viewer::inputAdvectedTexture[s][i][c];
			tap::f64(s ? "advected1" : "advected0", flat.data(), n, 3);
		}
		for (int i = 0; i < n; i++) Viewer::inputAdvectedTexture[0][i] = (Viewer::inputAdvectedTexture[0][i] + Viewer::inputAdvectedTexture[1][i]) / 2.0;
		OutputImage(Out.value, Viewer::inputAdvectedTexture[0], td.tWidth, td.tHeight, true);
	}
	else
	{
		Viewer::inputGeometryData.flow(fd, 0.5, Viewer::inputAdvectedSignal, Threads.value);
		tap::points3("advected0", Viewer::inputAdvectedSignal[0]), tap::points3("advected1", Viewer::inputAdvectedSignal[1]);
		int vCount = (int)fd.vertices.size();
		std::vector< Point3D< double > > outputColors(vCount);
		for (int v = 0; v < vCount; v++) outputColors[v] = Point3D< float >((Viewer::inputAdvectedSignal[0][v] + Viewer::inputAdvectedSignal[1][v]) / double(2.0));
		OutputMesh(Out.value, fd.vertices, outputColors, fd.triangles, PLY_ASCII);
	}
	return EXIT_SUCCESS;
}
